"""
I/O side of the export stage with the interface of the reference (sparseSpatialSampling/data.py):

* ``Datawriter(file_path, file_name, mode, mixed)`` -- ``write_data(name, data, group, time_step)``, ``write_grid``,
  ``write_xdmf_file``, ``close``, ``mode``, ``n_cells`` (data.py:303-501),
* ``Dataloader(load_path, file_name, dtype)`` -- ``write_times``, ``weights``, ``vertices``, ``nodes``, ``faces``,
  ``field_names``, ``levels``, ``metric``, ``load_snapshot`` (data.py:22-300),
* ``XDMFWriter(file_path, file_name, grid_name, mixed).write_xdmf()`` (data.py:504-777),

and the same on-disk layout: ``grid/{centers,vertices,faces}``, ``constant/*``, ``data/<time>/<field>_{center,vertices}``
(const.py:6-17, export.py:252-299).

File I/O is outside the accelerated path: this module is plain host Python. The files are real
HDF5: written through h5py when it can be imported, otherwise through ``h5lite`` -- a small pure-Python implementation of
the subset of the HDF5 format that h5py's defaults produce (version-0 superblock, symbol-table groups, contiguous
datasets), which also reads the reference's own output files. Data is appended to the file as it arrives (batch-wise
exports cost I/O linear in their size, nothing is kept in memory).
"""
import logging
from os.path import join, isfile
from typing import List, Union

import numpy as np
import torch as pt

from . import h5lite
from .const import CONST, GRID, DATA, CENTERS, VERTICES, FACES

logger = logging.getLogger(__name__)

try:                                    # pragma: no cover - depends on the image
    import h5py
    HAVE_H5PY = True
except ImportError:                     # pragma: no cover
    h5py = None
    HAVE_H5PY = False
FORCE_H5LITE = False                    # tests: exercise the built-in writer even when h5py is installed


def _to_numpy(data):
    if isinstance(data, pt.Tensor):
        return data.detach().cpu().numpy()
    return np.asarray(data)


class _H5Store:                         # pragma: no cover - needs h5py
    """The same small API on top of an HDF5 file."""

    def __init__(self, path: str, mode: str):
        self._file = h5py.File(path, mode)

    def keys(self, path: str = "") -> List[str]:
        node = self._file[path] if path else self._file
        return list(node.keys())

    def has(self, path: str) -> bool:
        return path in self._file

    def read(self, path: str) -> np.ndarray:
        return self._file[path][()]

    def shape(self, path: str) -> tuple:
        return tuple(self._file[path].shape)

    def write(self, group: str, name: str, data) -> bool:
        grp = self._file.require_group(group)
        if name in grp:
            return False
        grp.create_dataset(name, data=_to_numpy(data))
        return True

    def close(self) -> None:
        self._file.close()


class _LiteStore:
    """The same small API on top of ``h5lite.File`` (tensors are converted to numpy on the way in)."""

    def __init__(self, path: str, mode: str):
        self._file = h5lite.File(path, mode)
        self.keys, self.has, self.read, self.shape = self._file.keys, self._file.has, self._file.read, self._file.shape

    def write(self, group: str, name: str, data) -> bool:
        return self._file.write(group, name, _to_numpy(data))

    def close(self) -> None:
        self._file.close()


PART_FILES = "part_files"               # constant/part_files: newline-joined names of the other parts (UTF-8 bytes)


def _read_part_files(store) -> List[str]:
    """Names of the additional part files of a sharded export (``ExportData(distributed=True)``), [] otherwise."""
    if not store.has(f"{CONST}/{PART_FILES}"):
        return []
    raw = bytes(np.asarray(store.read(f"{CONST}/{PART_FILES}"), dtype=np.uint8))
    return [n for n in raw.decode("utf-8").split("\n") if n]


def _open_store(file_path: str, file_name: str, mode: str):
    full = join(file_path, file_name)
    if HAVE_H5PY and not FORCE_H5LITE:  # pragma: no cover
        return _H5Store(full, mode)
    return _LiteStore(full, mode)


# ---------------------------------------------------------------------------------------------------- Dataloader
class Dataloader:
    """Reads an S^3 output file (data.py:22-300)."""

    def __init__(self, load_path: str, file_name: str, dtype: pt.dtype = pt.float32):
        self._load_path = load_path
        self._file_name = file_name
        self._dtype = dtype
        self._size_initial_cell = None
        self._reset()

    def _store(self):
        return _open_store(self._load_path, self._file_name, "r")

    def _stores_by_time(self) -> dict:
        """time step -> store that holds it (the file itself, then the parts of a sharded export in rank order)."""
        if self._time_index is None:
            main = self._store()
            index = {}
            for st in [main] + [_open_store(self._load_path, name, "r") for name in _read_part_files(main)]:
                if st.has(DATA):
                    for t in st.keys(DATA):
                        index.setdefault(t, st)
            self._time_index = index
        return self._time_index

    def _reset(self) -> None:
        st = self._store()
        centers = st.read(f"{GRID}/{CENTERS}")
        self._n_cells, self._n_dimensions = centers.shape[0], centers.shape[1]
        try:
            self._size_initial_cell = st.read(f"{CONST}/size_initial_cell")
        except KeyError:
            logger.warning("Could not load initial cell size.")
        self._write_times = None
        self._time_index = None
        self._weights = None
        self._levels = None
        self._metric = None
        self._field_names = None
        self._vertices = None
        self._faces = None
        self._nodes = None

    def _read_tensor(self, path: str) -> pt.Tensor:
        return pt.from_numpy(np.array(self._store().read(path)))

    @property
    def write_times(self) -> List[str]:
        if self._write_times is None:
            index = self._stores_by_time()
            if index:
                self._write_times = sorted(index.keys(), key=lambda s: s.encode("utf-8"))   # HDF5 link order
        return self._write_times

    @property
    def weights(self) -> pt.Tensor:
        """cell areas (2D) / volumes (3D): (size_initial_cell / 2^level)^d  (data.py:240-247)"""
        if self._weights is None:
            size0 = float(np.asarray(self._size_initial_cell))
            self._weights = pt.pow(size0 / pt.pow(2.0, self.levels.to(pt.float64)), self._n_dimensions).squeeze()
        return self._weights

    @property
    def vertices(self) -> pt.Tensor:
        """cell centres (the reference calls them ``vertices`` here, data.py:95-108)"""
        if self._vertices is None:
            self._vertices = self._read_tensor(f"{GRID}/{CENTERS}")
        return self._vertices

    @property
    def nodes(self) -> pt.Tensor:
        if self._nodes is None:
            self._nodes = self._read_tensor(f"{GRID}/{VERTICES}")
        return self._nodes

    @property
    def faces(self) -> pt.Tensor:
        if self._faces is None:
            self._faces = self._read_tensor(f"{GRID}/{FACES}")
        return self._faces

    @property
    def field_names(self) -> dict:
        if self._field_names is None:
            index = self._stores_by_time()
            self._field_names = {t: [f.split("_")[0] for f in index[t].keys(f"{DATA}/{t}") if f.endswith("center")]
                                 for t in (self.write_times or [])}
        return self._field_names

    @property
    def levels(self) -> pt.Tensor:
        if self._levels is None:
            self._levels = self._read_tensor(f"{CONST}/levels").squeeze()
        return self._levels

    @property
    def metric(self) -> pt.Tensor:
        if self._metric is None:
            self._metric = self._read_tensor(f"{CONST}/metric").squeeze()
        return self._metric

    @property
    def load_path(self) -> str:
        return self._load_path

    @load_path.setter
    def load_path(self, value: str) -> None:
        self._load_path = value
        self._reset()

    @property
    def file_name(self) -> str:
        return self._file_name

    @file_name.setter
    def file_name(self, value: str) -> None:
        self._file_name = value
        self._reset()

    def load_snapshot(self, field_name: Union[List[str], str],
                      write_times: Union[str, List[str]] = None) -> Union[List[pt.Tensor], pt.Tensor]:
        """Data matrix ``[N_cells, T]`` (scalar) or ``[N_cells, D, T]`` (vector) of the fields at the cell centres."""
        if write_times is None:
            write_times = self.write_times
        if isinstance(write_times, str):
            write_times = [write_times]
        if isinstance(field_name, str):
            field_name = [field_name]
        index = self._stores_by_time()
        matrices = []
        for f in field_name:
            first = index[write_times[0]].read(f"{DATA}/{write_times[0]}/{f}_center")
            dm = pt.zeros(tuple(first.shape) + (len(write_times),), dtype=self._dtype)
            for i, t in enumerate(write_times):
                dm[..., i] = pt.from_numpy(np.array(index[t].read(f"{DATA}/{t}/{f}_center")))
            matrices.append(dm)
        return matrices[0] if len(matrices) == 1 else matrices


# ---------------------------------------------------------------------------------------------------- Datawriter
class Datawriter:
    def __init__(self, file_path: str, file_name: str, mode: str = "w", mixed: bool = False):
        self._file_name = file_name
        self._file_path = file_path
        self._mode = mode
        self._mixed = mixed
        self._n_cells = None
        self._closed = False
        self._store = _open_store(file_path, file_name, mode)

    # ------------------------------------------------------------------ reference interface
    def write_data(self, name: str, data: any, group: str = CONST, time_step: Union[int, float, str] = None) -> None:
        if group == DATA and time_step is None:
            logger.warning(f"No time step for group 'data' provided. Writing data to the zeroth time step '{DATA}/0'.")
            time_step = "0"
        if time_step is not None or group == DATA:
            if self._n_cells is not None and not (name.endswith("center") or name.endswith("vertices")):
                name = f"{name}_center" if data.shape[0] == self._n_cells else f"{name}_vertices"
            path = f"{DATA}/{time_step}"
        elif group == CONST:
            path = CONST
        elif group == GRID:
            path = GRID
        else:
            raise ValueError(f"Unknown group type, available types are '{DATA}', '{CONST}' and '{GRID}'.")
        if not self._store.write(path, name, data):
            logger.warning(f"Field {name} already exists in {path}. Skipping field {name}.")

    def write_grid(self, loader) -> None:
        self._n_cells = loader.vertices.shape[0]
        self.write_data(CENTERS, group=GRID, data=loader.vertices)
        self.write_data(VERTICES, group=GRID, data=loader.nodes)
        self.write_data(FACES, group=GRID, data=loader.faces)

    def write_part_files(self, names: List[str]) -> None:
        """Record the other part files of a sharded export (``constant/part_files``, newline-joined UTF-8 bytes)."""
        raw = np.frombuffer("\n".join(names).encode("utf-8"), dtype=np.uint8)
        self._store.write(CONST, PART_FILES, raw)

    def write_xdmf_file(self) -> None:
        logger.info(f"Writing XDMF file for file {self._file_name}")
        self.close()                    # the XDMF writer reads the finished file
        XDMFWriter(self._file_path, self._file_name, mixed=self._mixed).write_xdmf()

    def close(self) -> None:
        if not self._closed:
            self._store.close()
            self._closed = True

    @property
    def mode(self) -> str:
        return self._mode

    @mode.setter
    def mode(self, value) -> None:
        self.close()
        self._mode = value
        self._store = _open_store(self._file_path, self._file_name, value)
        self._closed = False

    @property
    def file_name(self) -> str:
        return self._file_name

    @property
    def n_cells(self):
        return self._n_cells

    @n_cells.setter
    def n_cells(self, value: int) -> None:
        self._n_cells = value


# ---------------------------------------------------------------------------------------------------- XDMFWriter
class XDMFWriter:
    """
    XDMF (version 2) description of an S^3 file for ParaView (data.py:504-777): one ``Uniform`` grid when the file has no
    temporal data, otherwise a temporal collection with one grid per time step; fields of ``constant`` whose first
    dimension matches the number of cells / vertices become attributes (of the first time step in the temporal case).
    """

    def __init__(self, file_path: str, file_name: str, grid_name: str = "grid_s_cube", mixed: bool = False):
        self._file_path = file_path
        self._grid_name = grid_name
        self._mixed = mixed
        self._hdf_file_name = file_name
        self._xdmf_file_name = f"{file_name.split('.h5')[0]}.xdmf"
        self._store = _open_store(file_path, file_name, "r")
        # sharded export: time steps live in the file itself or in one of its parts
        self._time_files = {}
        for name in [file_name] + _read_part_files(self._store):
            st = self._store if name == file_name else _open_store(file_path, name, "r")
            if st.has(DATA):
                for t in st.keys(DATA):
                    self._time_files.setdefault(t, (name, st))
        self._check_grid()
        centers = self._store.shape(f"{GRID}/{CENTERS}")
        self._n_dimensions = centers[-1]
        self._n_cells = centers[0]
        self._n_faces = self._store.shape(f"{GRID}/{FACES}")[0]
        self._n_vertices = self._store.shape(f"{GRID}/{VERTICES}")[0]
        if mixed:
            self._grid_type = "Mixed"
        else:
            self._grid_type = "Quadrilateral" if self._n_dimensions == 2 else "Hexahedron"
        self._dims = "XY" if self._n_dimensions == 2 else "XYZ"

    def _check_grid(self) -> None:
        if not self._store.has(GRID):
            raise RuntimeError("Found no grid in the provided HDF5 file. Unable to create XDMF file without a grid.")
        for key, what in ((FACES, "cell faces"), (CENTERS, "cell centers"), (VERTICES, "cell vertices")):
            if key not in self._store.keys(GRID):
                raise RuntimeError(f"Unable to find {what} in group {GRID}. Make sure the key is present and named "
                                   f"{key}.")

    # ------------------------------------------------------------------ building blocks
    def _topology_and_geometry(self) -> str:
        conn = f"{self._n_faces}" if self._mixed else f"{self._n_faces} {2 ** self._n_dimensions}"
        return (f'<Topology TopologyType="{self._grid_type}" NumberOfElements="{self._n_faces}">\n'
                f'<DataItem Format="HDF" DataType="Int" Dimensions="{conn}">\n'
                f'{self._hdf_file_name}:/{GRID}/{FACES}\n'
                f'</DataItem>\n</Topology>\n'
                f'<Geometry GeometryType="{self._dims}">\n'
                f'<DataItem Rank="2" Dimensions="{self._n_vertices} {self._n_dimensions}" NumberType="Float" '
                f'Precision="8" Format="HDF">\n'
                f'{self._hdf_file_name}:/{GRID}/{VERTICES}\n'
                f'</DataItem>\n</Geometry>\n')

    def _attribute(self, name: str, path: str, shape, hdf_file: str = None) -> str:
        """One attribute entry, or an empty string if the dataset lives neither on the cells nor on the vertices."""
        hdf_file = hdf_file or self._hdf_file_name
        if len(shape) == 0:
            return ""
        if shape[0] == self._n_cells:
            center, n = "Cell", self._n_cells
        elif shape[0] == self._n_vertices:
            center, n = "Node", self._n_vertices
        else:
            logger.warning(f"Field in '{path}' with a size of {tuple(shape)} doesn't match the number of cells with "
                           f"N_cells = {self._n_cells} or the number of vertices with N_vertices = "
                           f"{self._n_vertices}. Skipping this field.")
            return ""
        width = 1 if len(shape) == 1 else shape[1]
        return (f'<Attribute Name="{name}" AttributeType="Vector" Center="{center}">\n'
                f'<DataItem NumberType="Float" Precision="8" Format="HDF" Dimensions="{n} {width}">\n'
                f'{hdf_file}:/{path}\n</DataItem>\n</Attribute>\n')

    def _constant_attributes(self) -> str:
        if not self._store.has(CONST):
            logger.info("Couldn't find any constant fields to write.")
            return ""
        out = []
        for k in self._store.keys(CONST):
            if k == PART_FILES:
                continue
            shape = self._store.shape(f"{CONST}/{k}")
            if len(shape) and shape[0] in (self._n_cells, self._n_vertices):
                out.append(self._attribute(k, f"{CONST}/{k}", shape))
        return "".join(out)

    # ------------------------------------------------------------------ reference interface
    def write_xdmf(self) -> None:
        header = '<?xml version="1.0"?>\n<!DOCTYPE Xdmf SYSTEM "Xdmf.dtd" []>\n<Xdmf Version="2.0">\n'
        parts = [header]
        if self._time_files:
            parts.append(f'<Domain>\n<Grid Name="{self._grid_name}" GridType="Collection" CollectionType="temporal">\n')
            for i, t in enumerate(sorted(self._time_files.keys(), key=lambda x: float(x))):
                hdf_file, st = self._time_files[t]
                parts.append(f'<Grid Name="{self._grid_name} {t}" GridType="Uniform">\n<Time Value="{t}"/>\n')
                parts.append(self._topology_and_geometry())
                if i == 0:
                    parts.append(self._constant_attributes())
                for k in st.keys(f"{DATA}/{t}"):
                    # fields are stored as <field_name>_<position>
                    name = "_".join(k.split("_")[:-1]) if len(k.split("_")) > 1 else k
                    parts.append(self._attribute(name, f"{DATA}/{t}/{k}", st.shape(f"{DATA}/{t}/{k}"), hdf_file))
                parts.append('</Grid>\n')
            parts.append('</Grid>\n</Domain>\n</Xdmf>')
        else:
            parts.append(f'<Domain>\n<Grid Name="{self._grid_name}" GridType="Uniform">\n')
            parts.append(self._topology_and_geometry())
            parts.append(self._constant_attributes())
            parts.append("</Grid>\n</Domain>\n</Xdmf>")
        with open(join(self._file_path, self._xdmf_file_name), "w") as f_out:
            f_out.write("".join(parts))
