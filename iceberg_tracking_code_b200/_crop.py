"""Worker of Camera.crop_image_parallel, in a module of its own so that the spawned pool processes import Pillow only
(not torch / libibt.so)."""


def crop_image_standalone(args):
    """camtools.py:63-104: crop one photo with the given box and save it; truncated source files are cropped with
    ImageFile.LOAD_TRUNCATED_IMAGES like the reference does after its first attempt fails."""
    from PIL import Image, ImageFile
    inpath, outpath, box = args
    try:
        ImageFile.LOAD_TRUNCATED_IMAGES = False
        Image.open(inpath).crop(box).save(outpath)
    except Exception:                                      # noqa: BLE001  (the reference catches everything, camtools.py:82)
        ImageFile.LOAD_TRUNCATED_IMAGES = True
        Image.open(inpath).crop(box).save(outpath)
        print(str(inpath) + ' TRUNCATED...')
