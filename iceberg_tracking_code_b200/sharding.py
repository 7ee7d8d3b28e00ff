"""Time-block sharding of the tracking loop across the GPUs of one box.

A *group* = one seed frame + track_len frame pairs (s1_lucaskanade_tracking.py:304-307, 362, 437-448): seeds come from
the seed frame alone, tracks never outlive the group, the .npz is named after the seed frame -- so groups are
independent units.  Rank r of R takes a contiguous block of groups (= a contiguous time block of frames plus ONE halo
frame shared with the next rank) and there is no data-path communication.  The only collective is the final gather
of the per-group track arrays to rank 0 (`gather_results`; NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
import numpy as np
import torch


def n_groups(n_frames, track_len, start=0):
    """Completed groups of a sequence: the loop of s1:307 saves at counter = T, 2T, ... <= n-1."""
    return max(0, (n_frames - start - 1) // int(track_len))


def shard_groups(total_groups, rank, world):
    """Contiguous block [g0, g0+n) of rank `rank`; the remainder goes to the first ranks."""
    base, rem = divmod(int(total_groups), int(world))
    n = base + (1 if rank < rem else 0)
    g0 = rank * base + min(rank, rem)
    return g0, n


def frame_range(g0, n, track_len, start=0):
    """Inclusive frame index range [first, last] a block of groups needs (last = halo / next block's seed frame)."""
    if n <= 0:
        return None
    return start + g0 * track_len, start + (g0 + n) * track_len


def pack_results(results, track_len):
    """[(seed_index, path, tracks (M,T+1,2) f32 | (0,) f64, quality)] -> (meta int64 (G,2), tracks f32 (SM,T+1,2),
    quality f32 (SM,T))."""
    T = int(track_len)
    meta = np.zeros((len(results), 2), np.int64)
    tr, qu = [], []
    for i, (seed, _path, tracks, quality) in enumerate(results):
        m = 0 if tracks.ndim != 3 else tracks.shape[0]
        meta[i] = (seed, m)
        if m:
            tr.append(np.asarray(tracks, np.float32).reshape(m, T + 1, 2))
            qu.append(np.asarray(quality, np.float32).reshape(m, T))
    tracks = np.concatenate(tr, 0) if tr else np.zeros((0, T + 1, 2), np.float32)
    quality = np.concatenate(qu, 0) if qu else np.zeros((0, T), np.float32)
    return meta, tracks, quality


def unpack_results(meta, tracks, quality):
    """tracks / quality may be numpy arrays or torch tensors (device-resident gather): groups are views into them."""
    out, o = [], 0
    for seed, m in (meta.tolist() if hasattr(meta, "tolist") else meta):
        if m:
            out.append((int(seed), tracks[o:o + m], quality[o:o + m]))
        else:
            out.append((int(seed), np.zeros((0,), np.float64), np.zeros((0,), np.float64)))
        o += m
    return out


_pinned_cache = {}


def _pinned(shape, dtype):
    """pinned host buffers reused between gathers (cudaHostAlloc of ~100 MB costs tens of ms); one per (trailing shape,
    dtype), grown when a gather needs more rows"""
    key = (tuple(shape[1:]), dtype)
    t = _pinned_cache.get(key)
    if t is None or t.shape[0] < shape[0]:
        t = torch.empty(tuple(shape), dtype=dtype).pin_memory()
        _pinned_cache[key] = t
    return t[:shape[0]]


def gather_results(results, track_len, device=None, to_host=True, device_results=None):
    """The single collective of the path: every rank contributes its groups' track arrays; returns, on EVERY rank,
    the list [(seed_index, tracks, trackquality)] of all ranks in time order (all_gather of sizes, then all_gather of
    the padded payloads).  Without an initialised process group this is the identity.
    to_host=False leaves the gathered payloads on the device (the groups are views into one CUDA tensor per array):
    copying 8 ranks' worth of a day (0.6 GB) into fresh host memory costs more than tracking the day.
    to_host="rank0" is the single-writer mode: rank 0 copies the gathered arrays to (reused, pinned) host memory once, the
    other ranks keep device views.
    device_results: {seed_index: (tracks, trackquality) CUDA tensors} as track_sequence(device_results=) leaves them: the
    payload is then assembled on the device instead of being uploaded from the host arrays again."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [(int(seed), tracks, quality) for seed, _path, tracks, quality in results]      # one rank: nothing to move
    world = dist.get_world_size()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    host = bool(to_host) if to_host != "rank0" else dist.get_rank() == 0
    T = int(track_len)
    use_dev = device_results is not None and device.type == "cuda"
    if use_dev:
        meta = np.zeros((len(results), 2), np.int64)
        tr, qu = [], []
        for i, (seed, _path, tracks_h, _q) in enumerate(results):
            m = 0 if tracks_h.ndim != 3 else tracks_h.shape[0]
            meta[i] = (seed, m)
            if m:
                t_d, q_d = device_results[seed]
                tr.append(t_d.reshape(m, T + 1, 2)); qu.append(q_d.reshape(m, T))
        tracks = torch.cat(tr, 0) if tr else torch.zeros((0, T + 1, 2), dtype=torch.float32, device=device)
        quality = torch.cat(qu, 0) if qu else torch.zeros((0, T), dtype=torch.float32, device=device)
    else:
        meta, tracks, quality = pack_results(results, track_len)
    sizes = torch.tensor([meta.shape[0], tracks.shape[0]], dtype=torch.int64, device=device)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes)
    all_sizes = torch.stack(all_sizes).cpu().numpy()
    gmax, mmax = int(all_sizes[:, 0].max()), int(all_sizes[:, 1].max())

    def padded(a, n, dtype):
        t = torch.empty((n,) + tuple(a.shape[1:]), dtype=dtype, device=device)
        k = a.shape[0]
        if k:
            t[:k] = a if isinstance(a, torch.Tensor) else torch.from_numpy(a).to(device, non_blocking=True)
        if k < n:
            t[k:].zero_()
        return t

    def allg(t, host=True):
        """one collective into a tensor concatenated along dim 0; returns the per-rank slices"""
        n = t.shape[0]
        if n == 0:
            return [t.cpu().numpy() if host else t] * world
        out = torch.empty((world * n,) + tuple(t.shape[1:]), dtype=t.dtype, device=device)
        dist.all_gather_into_tensor(out, t)
        if host:
            if out.is_cuda:
                h = _pinned(out.shape, out.dtype)
                h.copy_(out, non_blocking=True)
                torch.cuda.current_stream().synchronize()
                out = h.numpy()
            else:
                out = out.numpy()
        return [out[r * n:(r + 1) * n] for r in range(world)]

    metas = allg(padded(meta, gmax, torch.int64), True) if gmax else [np.zeros((0, 2), np.int64)] * world
    trs = allg(padded(tracks, mmax, torch.float32), host) if mmax else [np.zeros((0, T + 1, 2), np.float32)] * world
    qus = allg(padded(quality, mmax, torch.float32), host) if mmax else [np.zeros((0, T), np.float32)] * world
    out = []
    for r in range(world):
        g, m = int(all_sizes[r, 0]), int(all_sizes[r, 1])
        out += unpack_results(metas[r][:g], trs[r][:m], qus[r][:m])
    out.sort(key=lambda x: x[0])
    return out


def track_sequence_sharded(imagelist, mask, track_len, track_len_sec, start=0, rank=None, world=None, gather=True, **kw):
    """Run this rank's block of groups of `imagelist` (tracking.track_sequence) and, if asked, gather all ranks'
    tracks.  Each rank writes its own .npz files (the file set is disjoint by construction)."""
    import torch.distributed as dist
    from .tracking import track_sequence
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    total = n_groups(len(imagelist), track_len, start)
    g0, n = shard_groups(total, rank, world)
    dev_res = {} if (gather and torch.cuda.is_available() and dist.is_initialized() and dist.get_backend() == "nccl") else None
    res = track_sequence(imagelist, mask, track_len, track_len_sec, startlist=(start,), first_group=g0, n_groups=n,
                         device_results=dev_res, **kw) if n else []
    if not gather:
        return [(s, t, q) for s, _p, t, q in res]
    return gather_results(res, track_len, device_results=dev_res)
