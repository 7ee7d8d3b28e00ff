"""Worker of Camera.crop_image_parallel, in a module of its own so that the spawned pool processes import Pillow only
(not torch / libibt.so)."""


def log_truncated(inpath, outpath):
    """camtools.py:84-92: the reference's log entry for a frame whose first read failed, in
    <two levels above the cropped frame>/logfile_image_cropping.log."""
    import datetime as dt
    import logging
    import os.path as osp
    logfilepath = osp.dirname(osp.dirname(str(outpath)))
    logging.basicConfig(level=logging.DEBUG, filename=logfilepath + '/logfile_image_cropping.log')
    logging.info('------------------')
    logging.info(inpath)
    logging.info(dt.datetime.now().isoformat())
    logging.exception('')


def open_cropped(inpath, outpath, box):
    """camtools.py:63-104 up to the save: the cropped PIL image; a truncated source is logged, read again with
    ImageFile.LOAD_TRUNCATED_IMAGES and announced like the reference does."""
    from PIL import Image, ImageFile
    try:
        ImageFile.LOAD_TRUNCATED_IMAGES = False
        img = Image.open(inpath).crop(box)
        img.load()
        return img
    except Exception:                                      # noqa: BLE001  (the reference catches everything, camtools.py:82)
        log_truncated(inpath, outpath)
        ImageFile.LOAD_TRUNCATED_IMAGES = True
        img = Image.open(inpath).crop(box)
        img.load()
        print(str(inpath) + ' TRUNCATED...')
        return img


def crop_image_standalone(args):
    """camtools.py:63-104: crop one photo with the given box and save it (Pillow's defaults, like the reference)."""
    inpath, outpath, box = args
    open_cropped(inpath, outpath, box).save(outpath)
