"""Stand-in for flowtorch.data (branch aweiner, not vendored): the two mask helpers the reference's geometries call."""
import torch as pt


def mask_box(vertices, lower, upper):
    mask = pt.ones(vertices.shape[0], dtype=pt.bool)
    for i, (lo, up) in enumerate(zip(lower, upper)):
        mask = mask & (vertices[:, i] >= lo) & (vertices[:, i] <= up)
    return mask


def mask_sphere(vertices, center, radius):
    loc = pt.tensor(center, dtype=vertices.dtype)
    return (vertices - loc).norm(dim=1) <= radius


class FOAMDataloader:
    pass
