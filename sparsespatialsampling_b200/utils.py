"""
Workflow helpers of the reference that sit on the accelerated path (sparseSpatialSampling/utils.py):
``compute_svd`` (utils.py:302-346) and ``write_svd_s_cube_to_file`` (utils.py:349-413).

The OpenFOAM loaders of the reference (``load_foam_data``, ``load_original_Foam_fields``, ``export_openfoam_fields``,
utils.py:23-299) are file I/O through flowtorch and out of scope.
"""
import logging
from typing import Union

from .data import Dataloader, Datawriter
from .svd import compute_svd

logger = logging.getLogger(__name__)

__all__ = ["compute_svd", "write_svd_s_cube_to_file"]


def write_svd_s_cube_to_file(field_names: Union[list, str], load_dir: str, file_name: str, new_file: bool,
                             n_modes: int = None, rank=None, t_start: Union[int, float] = 0,
                             method: str = "tc3") -> None:
    """
    SVD of exported fields, written as ``<file_name>_<field>_svd.h5`` (+ XDMF): modes ``mode_1..n`` , ``V``, ``s``,
    ``cell_area`` in ``constant`` (utils.py:349-413). Same arguments as the reference plus ``method`` (Gram kernel).
    Only the ``n_modes`` modes that are written are projected on the device; ``s`` and ``V`` keep ``rank`` entries.
    """
    if type(field_names) is str:
        field_names = [field_names]

    for f in field_names:
        logger.info(f"Performing SVD for field {f}.")
        _name = f"{file_name}_{f}" if new_file else file_name
        dataloader = Dataloader(load_dir, f"{_name}.h5")
        _write_times = sorted([t for t in dataloader.write_times if float(t) >= t_start], key=lambda x: float(x))

        # data matrix of the field, SVD weighted with the cell areas
        s, U, V = compute_svd(dataloader.load_snapshot(f, _write_times), dataloader.weights, rank, method=method,
                              n_modes=n_modes)

        datawriter = Datawriter(load_dir, file_name + f"_{f}_svd.h5")
        datawriter.write_grid(dataloader)

        n_write = U.size(-1) if n_modes is None else n_modes
        if n_write > U.size(-1):
            logger.warning(f"Number of modes to write is set to {n_write}, but found only {U.size(-1)} modes to write.")
            n_write = U.size(-1)

        # each mode is written as an independent field
        for i in range(n_write):
            datawriter.write_data(f"mode_{i + 1}", group="constant", data=U[..., i].squeeze())

        # not referenced by the XDMF file
        datawriter.write_data("V", group="constant", data=V)
        datawriter.write_data("s", group="constant", data=s)
        datawriter.write_data("cell_area", group="constant", data=dataloader.weights)
        datawriter.write_xdmf_file()
