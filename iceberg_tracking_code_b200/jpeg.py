"""`np.array(Image.open(image))` on the GPU (s1_lucaskanade_tracking.py:310, s0_1_test_lucaskanade_tracking.py:79).

The reference decodes every time-lapse frame with Pillow (libjpeg-turbo, ~0.25 s per 24 MP frame on one core) and
hands the (H,W,3) RGB array to cv2.cvtColor.  Here the compressed file goes to the GPU as it is (a few MB instead of
72 MB of pixels) and `ibt_jpeg_decode` (csrc/jpeg.cu) produces the same RGB array bit for bit -- and, fused, the gray
plane cv2.cvtColor would make of it, which is all the tracker consumes.

    imread(path_or_bytes)                 -> (H,W,3) u8 RGB CUDA tensor   == np.array(Image.open(path))
    imread(path_or_bytes, gray=True)      -> (H,W) u8 CUDA tensor         == cv2.cvtColor(that, cv2.COLOR_BGR2GRAY)
    JpegDecoder                           reusable pinned staging + device workspace (steady state allocates nothing)

Files the kernels do not handle (progressive, arithmetic coding, CMYK, 12-bit, exotic sampling) raise `Unsupported`;
`tracking.load_image` is the caller's way out (Pillow on the host, pixels uploaded) -- the compute path itself has no
CPU fallback.
"""
import ctypes as C

import numpy as np
import torch

from . import _native as N
from . import cv


SUBSAMPLING = {"4:4:4": (1, 1), "4:2:2": (2, 1), "4:2:0": (2, 2), 0: (1, 1), 1: (2, 1), 2: (2, 2)}     # Image.save(subsampling=)


class Unsupported(ValueError):
    """A valid JPEG this decoder does not handle (IBT_E_UNSUPPORTED)."""


def set_probe(enabled):
    """Process-wide (ibt_jpeg_set_probe): the entry-state probe of the Huffman pass shortens a lone decode; frame loops that
    keep several asynchronous decodes in flight beside the tracker switch it off (fewer full-grid kernels, ~4 % per step)."""
    N.check(N.lib().ibt_jpeg_set_probe(1 if enabled else 0), "ibt_jpeg_set_probe")


def parse(data):
    """Marker parsing on the host (ibt_jpeg_parse).  data: bytes-like.  Returns the filled ibt_jpeg_info_t."""
    buf = np.frombuffer(data, dtype=np.uint8)
    info = N.ibt_jpeg_info_t()
    rc = N.lib().ibt_jpeg_parse(C.c_void_p(buf.ctypes.data), buf.size, C.byref(info))
    if rc == N.IBT_E_UNSUPPORTED:
        raise Unsupported("JPEG variant not handled on the GPU (progressive / arithmetic / sampling / components)")
    N.check(rc, "ibt_jpeg_parse")
    return info


class JpegDecoder:
    """One per (GPU, decoding thread).  decode() is asynchronous on the current stream up to the convergence check of
    the speculative Huffman pass (one small host read per file)."""

    def __init__(self, device=None):
        self.device = cv._device() if device is None else torch.device(device)
        self._pinned = [None, None]
        self._k = 0
        self._dev_file = None
        self._ws = None
        self.last_rounds = 0

    def _pinned_slot(self, n):
        """the next pinned staging buffer (double-buffered), free again, at least n bytes"""
        k = self._k
        self._k ^= 1
        slot = self._pinned[k]
        if slot is None or slot[0].numel() < n:
            slot = (torch.empty((max(n, 1 << 20) * 5 // 4,), dtype=torch.uint8).pin_memory(), torch.cuda.Event())
            self._pinned[k] = slot
        slot[1].synchronize()                       # the copy that last read this buffer has finished
        return slot

    def _upload(self, slot, n):
        """pinned bytes [0, n) -> device, asynchronously on the current stream"""
        host, ev = slot
        if self._dev_file is None or self._dev_file.numel() < n:
            self._dev_file = torch.empty((max(n, 1 << 20) * 5 // 4,), dtype=torch.uint8, device=self.device)
        self._dev_file[:n].copy_(host[:n], non_blocking=True)
        ev.record(torch.cuda.current_stream())
        return self._dev_file

    def decode_file(self, path, rgb=True, gray=False, coeffset=0):
        """decode() of a file on disk.  The file is read into ordinary memory and copied into the pinned staging buffer:
        reading straight into page-locked memory (readinto on the pinned view) was measured 1.2x .. 30x slower and erratic."""
        with open(str(path), "rb") as f:
            return self.decode(f.read(), rgb, gray, coeffset)

    def decode(self, data, rgb=True, gray=False, coeffset=0):
        """data: bytes-like holding a JPEG file.  Returns (rgb, gray) CUDA tensors (None where not requested)."""
        buf = np.frombuffer(data, dtype=np.uint8)
        slot = self._pinned_slot(buf.size)
        view = slot[0].numpy()
        view[:buf.size] = buf
        return self._decode_staged(view[:buf.size], slot, rgb, gray, coeffset)

    def _decode_staged(self, buf, slot, rgb, gray, coeffset):
        info = parse(buf)
        H, W = info.height, info.width
        if info.ncomp == 1:
            rgb, gray = False, True                       # mode "L": np.array(Image.open(f)) is already (H,W)
        need = N.lib().ibt_jpeg_workspace_bytes(C.byref(info))
        if need <= 0:
            raise Unsupported("JPEG variant not handled on the GPU")
        with torch.cuda.device(self.device):
            if self._ws is None or self._ws.numel() < need:
                self._ws = torch.empty((int(need * 1.1) + 256,), dtype=torch.uint8, device=self.device)
            dfile = self._upload(slot, buf.size)
            out_rgb = torch.empty((H, W, 3), dtype=torch.uint8, device=self.device) if rgb else None
            out_gray = torch.empty((H, W), dtype=torch.uint8, device=self.device) if gray else None
            rounds = C.c_int(self.last_rounds)         # hint: consecutive frames of a sequence need about the same
            N.check(N.lib().ibt_jpeg_decode(cv._ptr(dfile), C.byref(info), cv._ptr(self._ws), self._ws.numel(),
                                            cv._ptr(out_rgb), W * 3, cv._ptr(out_gray), W, int(coeffset),
                                            C.byref(rounds), cv._stream()), "ibt_jpeg_decode")
            self.last_rounds = rounds.value
        return out_rgb, out_gray


    def recompress(self, rgb_in, quality=75, subsampling="4:2:0", rgb=True, gray=False, coeffset=0):
        """What np.array(Image.open(f)) holds after Image.fromarray(rgb_in).save(f, quality=, subsampling=) -- the
        save-and-reopen round trip of the reference's cropping pre-pass (imports/camtools.py:80,102,232 -> s1:310), bit-exact,
        without the file (ibt_jpeg_recompress).  rgb_in: (H,W,3) u8 CUDA tensor, may be a crop VIEW of a decoded frame (last two
        strides (3, 1)).  Pillow's defaults are the defaults.  Returns (rgb, gray) like decode(); asynchronous on the current stream."""
        hs, vs = SUBSAMPLING[subsampling]
        if rgb_in.dim() != 3 or rgb_in.shape[2] != 3 or rgb_in.dtype != torch.uint8 or not rgb_in.is_cuda:
            raise ValueError("recompress: (H,W,3) u8 CUDA tensor expected")
        if rgb_in.stride(2) != 1 or rgb_in.stride(1) != 3:
            rgb_in = rgb_in.contiguous()
        H, W = int(rgb_in.shape[0]), int(rgb_in.shape[1])
        need = N.lib().ibt_jpeg_recompress_workspace_bytes(W, H, hs, vs)
        if need <= 0:
            raise ValueError("recompress: image size / subsampling not handled")
        with torch.cuda.device(self.device):
            ws = getattr(self, "_ws_enc", None)
            if ws is None or ws.numel() < need:
                ws = self._ws_enc = torch.empty((int(need * 1.1) + 256,), dtype=torch.uint8, device=self.device)
            out_rgb = torch.empty((H, W, 3), dtype=torch.uint8, device=self.device) if rgb else None
            out_gray = torch.empty((H, W), dtype=torch.uint8, device=self.device) if gray else None
            N.check(N.lib().ibt_jpeg_recompress(cv._ptr(rgb_in), int(rgb_in.stride(0)), W, H, int(quality), hs, vs, cv._ptr(ws),
                                                ws.numel(), cv._ptr(out_rgb), W * 3, cv._ptr(out_gray), W, int(coeffset),
                                                cv._stream()), "ibt_jpeg_recompress")
        return out_rgb, out_gray

    # -- staging on a helper thread: file read, copy into page-locked memory and marker parsing cost ~2 ms of host time per
    # 24 MP frame -- more than the GPU needs for the frame -- so the frame loop runs them ahead on worker threads --------------
    NSTAGE = 4

    def stage(self, src):
        """src: path or bytes.  Reads the file, copies it into one of NSTAGE pinned staging buffers and parses its markers.
        Safe to call from worker threads (one decoder may be staged by several); the returned handle goes to
        decode_async(staged=...) on the thread that owns the CUDA stream, in the order the frames were staged."""
        import threading
        if not isinstance(src, (bytes, bytearray, memoryview, np.ndarray)):
            with open(str(src), "rb") as f:
                src = f.read()
        buf = np.frombuffer(src, dtype=np.uint8)
        if not hasattr(self, "_stage_lock"):
            self._stage_lock = threading.Lock()
        with self._stage_lock:
            if not hasattr(self, "_stage_slots"):
                self._stage_slots, self._stage_k = [None] * self.NSTAGE, 0
            k = self._stage_k
            self._stage_k = (k + 1) % self.NSTAGE
            slot = self._stage_slots[k]
            if slot is None:
                free = threading.Event(); free.set()
                slot = self._stage_slots[k] = {"host": None, "ev": torch.cuda.Event(), "free": free}
        slot["free"].wait()                               # the decode that used this buffer has been enqueued ...
        slot["ev"].synchronize()                          # ... and its host -> device copy has finished
        slot["free"].clear()
        if slot["host"] is None or slot["host"].numel() < buf.size:
            slot["host"] = torch.empty((max(buf.size, 1 << 20) * 5 // 4,), dtype=torch.uint8).pin_memory()
        view = slot["host"].numpy()
        view[:buf.size] = buf
        try:
            info = parse(view[:buf.size])
        except Exception:
            slot["free"].set()
            raise
        return {"slot": slot, "n": int(buf.size), "info": info, "data": src}

    # -- asynchronous form: nothing waits for the GPU; convergence of the speculative Huffman pass is checked afterwards --------
    def decode_async(self, data=None, rgb=False, gray=True, coeffset=0, margin=4, staged=None):
        """Like decode(), but the host does not wait: `last_rounds + margin` synchronisation rounds are enqueued (consecutive
        frames of a camera need about the same number; the first file of a decoder goes through decode()).  Returns a handle;
        confirm(handle) waits for the decode and returns (rgb, gray) -- of a repeated, synchronous decode if the Huffman pass
        had not converged within the rounds given.  staged: a handle of stage() instead of `data`."""
        if staged is None:
            staged = self.stage(data)
        slot, n, info, data = staged["slot"], staged["n"], staged["info"], staged["data"]
        host = slot["host"]
        try:
            if self.last_rounds <= 0:
                r, g = self._decode_staged(host.numpy()[:n], (host, slot["ev"]), rgb, gray, coeffset)
                ev = torch.cuda.Event(); ev.record()
                return {"rgb": r, "gray": g, "event": ev, "flags": None, "rounds": 0, "data": None}
            H, W = info.height, info.width
            if info.ncomp == 1:
                rgb, gray = False, True
            need = N.lib().ibt_jpeg_workspace_bytes(C.byref(info))
            if need <= 0:
                raise Unsupported("JPEG variant not handled on the GPU")
            rounds = min(64, self.last_rounds + int(margin))
            hb = int(N.lib().ibt_jpeg_async_host_bytes())
            if not hasattr(self, "_hpin"):
                self._hpin, self._hk = [None] * 4, 0
            k = self._hk
            self._hk = (k + 1) % 4
            hp = self._hpin[k]
            if hp is None:
                hp = self._hpin[k] = (torch.empty((hb,), dtype=torch.uint8).pin_memory(), torch.cuda.Event())
            hp[1].synchronize()                          # the decode that used this staging four calls ago has passed
            with torch.cuda.device(self.device):
                if self._ws is None or self._ws.numel() < need:
                    self._ws = torch.empty((int(need * 1.1) + 256,), dtype=torch.uint8, device=self.device)
                dfile = self._upload((host, slot["ev"]), n)
                out_rgb = torch.empty((H, W, 3), dtype=torch.uint8, device=self.device) if rgb else None
                out_gray = torch.empty((H, W), dtype=torch.uint8, device=self.device) if gray else None
                N.check(N.lib().ibt_jpeg_decode_async(cv._ptr(dfile), C.byref(info), cv._ptr(self._ws), self._ws.numel(),
                                                      cv._ptr(out_rgb), W * 3, cv._ptr(out_gray), W, int(coeffset), rounds,
                                                      C.c_void_p(hp[0].data_ptr()), hb, cv._stream()), "ibt_jpeg_decode_async")
                hp[1].record(torch.cuda.current_stream())
                ev = torch.cuda.Event(); ev.record()
            return {"rgb": out_rgb, "gray": out_gray, "event": ev, "flags": hp[0], "rounds": rounds, "data": bytes(data),
                    "args": (rgb, gray, coeffset)}
        finally:
            slot["free"].set()                            # (the CUDA event of the slot guards the copy itself)

    def confirm(self, h):
        """Wait for the decode of handle h and check that its Huffman pass had converged; if not (the file needed more rounds
        than its predecessors) the file is decoded again synchronously INTO THE SAME outputs' place: returns (rgb, gray)."""
        h["event"].synchronize()
        if h["flags"] is None:
            return h["rgb"], h["gray"]
        fl = h["flags"][:4 * h["rounds"]].numpy().view(np.uint32)
        zero = np.flatnonzero(fl == 0)
        if zero.size:
            self.last_rounds = int(zero[0]) + 1
            return h["rgb"], h["gray"]
        self.last_rounds = 0                              # not converged: decode() finds the count again
        h["redo"] = True                                  # (whatever the caller derived from the outputs must be rebuilt)
        rgb, gray, coeffset = h["args"]
        r, g = self.decode(h["data"], rgb, gray, coeffset)
        if h["rgb"] is not None:
            h["rgb"].copy_(r)
        if h["gray"] is not None:
            h["gray"].copy_(g)
        torch.cuda.current_stream().synchronize()
        return h["rgb"], h["gray"]


_default = {}


def _decoder(device=None):
    dev = cv._device() if device is None else torch.device(device)
    if dev not in _default:
        _default[dev] = JpegDecoder(dev)
    return _default[dev]


def imread(src, gray=False, device=None, coeffset=0):
    """src: path or bytes.  gray=False: (H,W,3) RGB like np.array(Image.open(src)); gray=True: the (H,W) plane
    cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY) makes of it (s1:310-311)."""
    if not isinstance(src, (bytes, bytearray, memoryview, np.ndarray)):
        with open(str(src), "rb") as f:
            src = f.read()
    r, g = _decoder(device).decode(src, rgb=not gray, gray=gray, coeffset=coeffset)
    return g if (gray or r is None) else r


def save_reopen(rgb, quality=75, subsampling="4:2:0", gray=False, device=None, coeffset=0):
    """np.array(Image.open(f)) after Image.fromarray(rgb).save(f) (the reference's crop_image_standalone + s1:310), computed
    on the GPU without the file.  rgb: (H,W,3) u8 CUDA tensor or numpy array.  gray=True: the cvtColor plane of the result."""
    dec = _decoder(device)
    if isinstance(rgb, np.ndarray):
        rgb = torch.from_numpy(np.ascontiguousarray(rgb)).to(dec.device)
    r, g = dec.recompress(rgb, quality, subsampling, rgb=not gray, gray=gray, coeffset=coeffset)
    return g if gray else r
