from . import atari_wrappers, tile_images  # noqa: F401
