"""Frame preprocessing in front of the VQGAN — host mirror of `Sampler._process_frame` (lwm/vision_chat.py:59-74; also
the resize/crop of lwm/vision_generation.py's input path): resize so that the SHORT side becomes `size` (PIL's default
resampling, aspect ratio kept, the long side truncated to int), crop the central size x size window, scale uint8
[0,255] to float32 [-1,1] (x / 127.5 - 1). The reference does this on the host with PIL; so does this mirror — the
result is the `pixel_values` tensor `VQGAN.encode` takes (vision_chat.py:89-100)."""
import numpy as np


def process_frame(image, size=256):
    """image: PIL.Image -> float32 array [size, size, C] in [-1, 1]"""
    width, height = image.size
    if width < height:
        new_size = (size, int(size * height / width))
    else:
        new_size = (int(size * width / height), size)
    image = image.resize(new_size)
    left, top = (new_size[0] - size) / 2, (new_size[1] - size) / 2
    image = image.crop((left, top, left + size, top + size))
    return np.array(image, dtype=np.float32) / 127.5 - 1


def process_frames(images, size=256):
    """a list of PIL images (the frames of a clip, or one still image) -> float32 [T, size, size, C]"""
    return np.stack([process_frame(im, size) for im in images])
