"""Screen-tile partition used for multi-GPU rendering: tiles of `tile` x `tile` pixels, tile t belongs to rank t % world,
pixels inside a tile in Morton order.  Pure-Python mirror of the table built in rtxpt_b200/csrc/api.cu (ensureTargets); used by the CPU
tests of the gather logic and to interpret the compact all-gather buffers on the host."""
import numpy as np


def _morton_xy(tile):
    m = np.arange(tile * tile, dtype=np.uint32)
    x = np.zeros_like(m); y = np.zeros_like(m)
    for b in range(16):
        x |= ((m >> (2 * b)) & 1) << b
        y |= ((m >> (2 * b + 1)) & 1) << b
    return x, y


def pixel_table(width, height, tile, rank, world):
    """Packed (x << 16) | y of every pixel rank `rank` renders, in path-slot order."""
    tiles_x, tiles_y = (width + tile - 1) // tile, (height + tile - 1) // tile
    mx, my = _morton_xy(tile)
    out = []
    for t in range(rank, tiles_x * tiles_y, world):
        tx, ty = (t % tiles_x) * tile, (t // tiles_x) * tile
        x, y = tx + mx, ty + my
        keep = (x < width) & (y < height)
        out.append(((x[keep] << 16) | y[keep]).astype(np.uint32))
    return np.concatenate(out) if out else np.zeros(0, np.uint32)


def gather_layout(width, height, tile, world):
    """(tables per rank, padded length): layout of the all-gather buffer, padding entries are 0xFFFFFFFF."""
    tables = [pixel_table(width, height, tile, r, world) for r in range(world)]
    padded = max(len(t) for t in tables)
    return tables, padded


def pack_owned(image, table, padded):
    out = np.zeros((padded, image.shape[-1]), image.dtype)
    x, y = table >> 16, table & 0xFFFF
    out[:len(table)] = image[y, x]
    return out


def unpack_all(gathered, tables, padded, image):
    for r, table in enumerate(tables):
        x, y = table >> 16, table & 0xFFFF
        image[y, x] = gathered[r * padded:r * padded + len(table)]
    return image


# ---- generic exchange of several per-pixel images (rtxpt_b200_exchange_pack / _unpack): a rank's block holds, image after image, `padded` elements in slot order, every
# segment starting on a 16-byte boundary
def exchange_layout(bytes_per_pixel, padded):
    """(segment offsets, bytes per rank) for images of the given element sizes."""
    offs, off = [], 0
    for b in bytes_per_pixel:
        offs.append(off); off += (padded * b + 15) & ~15
    return offs, off


def exchange_pack(images, table, padded):
    """images: list of H x W (x C) arrays; returns this rank's block as uint8."""
    sizes = [int(np.prod(im.shape[2:], dtype=np.int64)) * im.dtype.itemsize for im in images]
    offs, total = exchange_layout(sizes, padded)
    out = np.zeros(total, np.uint8); x, y = table >> 16, table & 0xFFFF
    for im, b, o in zip(images, sizes, offs):
        seg = np.ascontiguousarray(im[y, x]).view(np.uint8).reshape(-1)
        out[o:o + len(seg)] = seg
    return out


def exchange_unpack(gathered, tables, padded, images, skip_rank=None):
    """gathered: world x bytes-per-rank uint8; scatters every rank's pixels (but skip_rank's) into `images` in place."""
    sizes = [int(np.prod(im.shape[2:], dtype=np.int64)) * im.dtype.itemsize for im in images]
    offs, total = exchange_layout(sizes, padded); g = np.asarray(gathered, np.uint8).reshape(len(tables), total)
    for r, table in enumerate(tables):
        if r == skip_rank: continue
        x, y = table >> 16, table & 0xFFFF
        for im, b, o in zip(images, sizes, offs):
            seg = g[r, o:o + len(table) * b]
            im[y, x] = seg.view(im.dtype).reshape((len(table),) + im.shape[2:])
    return images
