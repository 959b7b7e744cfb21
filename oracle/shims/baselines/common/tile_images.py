def tile_images(imgs):
    return imgs[0]
