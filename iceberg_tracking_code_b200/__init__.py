"""B200-native tracking hot path of glacierbliss/iceberg_tracking_code.

    from iceberg_tracking_code_b200 import cv          # cvtColor, goodFeaturesToTrack, calcOpticalFlowPyrLK, ...
    from iceberg_tracking_code_b200.tracking import lucaskanade_tracking, LucasKanade, SequenceTracker

All compute goes through libibt.so (include/ibt.h); there is no CPU fallback.
"""
__version__ = "0.1.0"
