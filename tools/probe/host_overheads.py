import io, os, sys, time, tempfile
sys.path.insert(0, ".")
import numpy as np, torch
from PIL import Image
from iceberg_tracking_code_b200 import build, synthetic as syn, jpeg
from iceberg_tracking_code_b200.tracking import SequenceTracker, read_file
build.build()
H, W = 4000, 6000
base = syn.base_texture(H, W, 7, device="cuda")
tmp = tempfile.mkdtemp()
files = []
for t in range(4):
    f = os.path.join(tmp, "f%d.jpg" % t); Image.fromarray(syn.frame_rgb(base, t, seed=7).cpu().numpy()).save(f); files.append(f)
dec = jpeg.JpegDecoder()
def T(fn, n=10):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): r = fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
data = read_file(files[0])
print("read_file %.3f ms" % T(lambda: read_file(files[1])))
print("parse %.3f ms" % T(lambda: jpeg.parse(data)))
buf = np.frombuffer(data, np.uint8)
print("stage %.3f ms" % T(lambda: dec._stage(buf)))
print("decode total %.3f ms" % T(lambda: dec.decode(data, rgb=False, gray=True)))
gp = dict(maxCorners=20000, qualityLevel=0.007, minDistance=10, blockSize=10)
lp = dict(winSize=(31, 31), maxLevel=4, criteria=(3, 30, 0.01))
trk = SequenceTracker(gp, lp)
g = [dec.decode(read_file(f), rgb=False, gray=True)[1].clone() for f in files]
print("prepare %.3f ms" % T(lambda: trk.prepare(g[0])))
p = [trk.prepare(x) for x in g]
print("seed %.3f ms" % T(lambda: trk.seed(p[0], None, 2)))
def tr():
    trk.seed(p[0], None, 2, points=pts); trk.track(p[0], p[1])
trk.seed(p[0], None, 2); pts = trk._tracks[0].clone().reshape(-1,1,2)
print("seed(points)+track %.3f ms" % T(tr))
def hv():
    trk.seed(p[0], None, 2, points=pts); trk.track(p[0], p[1]); trk.track(p[1], p[2]); return trk.harvest()
print("seed+2 tracks+harvest %.3f ms" % T(hv))
