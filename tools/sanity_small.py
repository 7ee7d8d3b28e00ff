"""Small end-to-end exercise of every kernel (for compute-sanitizer): odd sizes, borders, masks, all LK variants."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from iceberg_tracking_code_b200 import cv, synthetic as syn
from iceberg_tracking_code_b200.tracking import SequenceTracker
from iceberg_tracking_code_b200.camera import Camera

for (h, w) in [(123, 187), (64, 48), (201, 333)]:
    base = syn.base_texture(h, w, 3)
    rgb = [syn.frame_rgb(base, t, seed=3).numpy() for t in range(3)]
    g = [cv.cvtColor(f, cv.COLOR_BGR2GRAY) for f in rgb]
    cv.pyrDown(g[0]); cv.buildOpticalFlowPyramid(g[0], (21, 21), 3, True)
    mask = np.zeros((h, w), np.uint8); mask[5:h - 7, 9:w - 3] = 255
    for gp in (dict(maxCorners=0, qualityLevel=0.01, minDistance=7, blockSize=3), dict(maxCorners=50, qualityLevel=0.01, minDistance=0, blockSize=10)):
        p = cv.goodFeaturesToTrack(g[0], mask=mask, **gp)
    pts = np.concatenate([p.reshape(-1, 2), np.float32([[0, 0], [w - 1, h - 1], [-3, 5], [w + 2, h / 2]])]).reshape(-1, 1, 2)
    for win, ml in (((21, 21), 3), ((31, 31), 4), ((35, 35), 4), ((15, 9), 2), ((41, 41), 5)):
        cv.calcOpticalFlowPyrLK(g[0], g[1], pts, None, winSize=win, maxLevel=ml, criteria=(3, 30, 0.01))
        cv.calcOpticalFlowPyrLK_FB(g[0], g[1], pts, winSize=win, maxLevel=ml)
    trk = SequenceTracker(dict(maxCorners=200, qualityLevel=0.01, minDistance=5, blockSize=3), dict(winSize=(21, 21), maxLevel=3, criteria=(3, 30, 0.01)))
    pyr = [trk.prepare(f) for f in rgb]
    trk.seed(pyr[0], mask, 2); trk.track(pyr[0], pyr[1]); trk.track(pyr[1], pyr[2]); trk.harvest()
cam = Camera("c", parameters=dict(image_width=640, image_height=480, sensor_width=22.3, easting=1.0, northing=2.0, elevation=100.0,
                                  antenna_height=1.0, theta=300.0, phi=5.0, psi=-1.0, sigma=18.0, crop_left=3, crop_top=4),
             maskpoly=[(10, 10), (300, 20), (280, 200), (20, 180)])
cam.mask_image(211, 317); cam.tracks_to_utm(np.float32(np.random.rand(10, 3, 2) * 100 + 200))
torch.cuda.synchronize(); print("sanity ok")
