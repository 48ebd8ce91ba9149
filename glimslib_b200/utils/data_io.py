"""Mesh / function HDF5 I/O of the drop-in: the subset of ``glimslib/utils/data_io.py`` that sits
directly before and after the hot path -- ``save_mesh_hdf5`` / ``read_mesh_hdf5`` (:663-713, how the MPI examples load
pre-partitioned meshes, README.md:162-183), ``save_functions_hdf5`` / ``read_function_hdf5`` (:715-760).
VTU import / export and the post-run VTU merge (:423-654) are provided on the package's own VTU parser
(``backend/vtu.py``; vtk and meshio are not dependencies here): ``read_vtk_convert_to_fenics``,
``convert_meshio_to_fenics_mesh`` / ``convert_fenics_mesh_to_meshio`` (on a meshio-shaped container),
``identify_orphaned_vertices`` / ``remove_orphaned_vertices``, ``remove_mesh_subdomain``, ``merge_vtus_timestep`` /
``merge_VTUs``.  Image (SimpleITK) and ANTs utilities stay out of scope."""
import os

import numpy as np

from glimslib_b200 import config
from glimslib_b200 import fenics_local as fenics
from glimslib_b200.backend import vtu as _vtu
from glimslib_b200.utils import file_utils as fu


# ==============================================================================
# VTU / meshio-shaped meshes <-> backend meshes (data_io.py:423-585)
# ==============================================================================
def identify_orphaned_vertices(mesh_in):
    """Vertex ids that no cell references (data_io.py:423-432)."""
    used = np.zeros(mesh_in.num_vertices(), dtype=bool)
    used[np.asarray(mesh_in.cells).ravel()] = True
    return list(np.nonzero(~used)[0])


def remove_orphaned_vertices(mesh_in, vertex_ids):
    """New mesh without the given (unreferenced) vertices, connectivity renumbered (data_io.py:434-468)."""
    keep = np.ones(mesh_in.num_vertices(), dtype=bool)
    keep[np.asarray(vertex_ids, dtype=np.int64)] = False
    new_id = np.cumsum(keep) - 1
    return fenics.Mesh(mesh_in.coords[keep], new_id[np.asarray(mesh_in.cells)].astype(np.int32))


def convert_meshio_to_fenics_mesh(meshio_mesh, domain_array_name='ElementBlockIds'):
    """meshio-shaped mesh (``points``, ``cells{type: conn}``, ``cell_data{type: {name: array}}``) -> (mesh, subdomains);
    2D meshes stored with a zero third coordinate lose it, orphaned vertices are removed (data_io.py:470-524)."""
    cell_type = list(meshio_mesh.cells.keys())[0]
    if cell_type not in ('triangle', 'tetrahedron'):
        raise ValueError("Do not understand cell type '%s'" % cell_type)
    dim = 2 if cell_type == 'triangle' else 3
    cells = np.asarray(meshio_mesh.cells[cell_type])
    points = np.asarray(meshio_mesh.points, dtype=np.float64)
    if dim == 2 and points.shape[1] == 3:
        if np.all(points[:, 2] == 0):
            points = points[:, :2]
        else:
            raise ValueError("expect third coordinate of all points of a 2D mesh to be 0")
    mesh = fenics.Mesh(points[:, :dim], cells.astype(np.int32))
    orphans = identify_orphaned_vertices(mesh)
    new_mesh = remove_orphaned_vertices(mesh, orphans) if len(orphans) > 0 else mesh
    subdomains = fenics.MeshFunction("size_t", new_mesh, new_mesh.geometry().dim())
    subdomains.set_all(0)
    material = meshio_mesh.cell_data.get(cell_type, {}).get(domain_array_name)
    if material is not None:
        subdomains.array()[:] = np.asarray(material).astype(np.uint64)
    return new_mesh, subdomains


def convert_fenics_mesh_to_meshio(fenics_mesh, subdomains=None):
    """(data_io.py:527-546)"""
    dim = fenics_mesh.geometry().dim()
    cell_type = 'triangle' if dim == 2 else 'tetrahedron'
    mio_mesh = _vtu.VtuMesh(fenics_mesh.coordinates(), {cell_type: np.asarray(fenics_mesh.cells)})
    if subdomains is not None:
        mio_mesh.cell_data[cell_type] = {'ElementBlockIds': np.asarray(subdomains.array()).astype(np.int64)}
    return mio_mesh


def read_vtk_convert_to_fenics(path_to_vtk):
    """VTU file -> (mesh, subdomains); the first cell-data array is the tissue label, as in
    ``convert_vtk_mesh_to_meshio`` (data_io.py:549-577)."""
    m = _vtu.read_vtu(path_to_vtk)
    for ct, arrays in list(m.cell_data.items()):
        if arrays and 'ElementBlockIds' not in arrays:
            arrays['ElementBlockIds'] = next(iter(arrays.values()))
    return convert_meshio_to_fenics_mesh(m)


def remove_mesh_subdomain(fenics_mesh, subdomains, lower_thr, upper_thr, temp_dir=config.output_dir_temp):
    """Mesh restricted to the cells whose label lies in [lower_thr, upper_thr] (data_io.py:579-600; the reference
    round-trips through a vtk threshold filter, the selection is done directly here)."""
    lab = np.asarray(subdomains.array())
    sel = (lab >= lower_thr) & (lab <= upper_thr)
    cell_type = 'triangle' if fenics_mesh.geometry().dim() == 2 else 'tetrahedron'
    mio = _vtu.VtuMesh(fenics_mesh.coordinates(), {cell_type: np.asarray(fenics_mesh.cells)[sel]},
                       cell_data={cell_type: {'ElementBlockIds': lab[sel].astype(np.int64)}})
    return convert_meshio_to_fenics_mesh(mio)


# ==============================================================================
# POSTPROCESSING VTU OUTPUT (data_io.py:606-654)
# ==============================================================================
def create_file_name(name, step):
    return "%s_%05d000000.vtu" % (name, step)


def remove_vtu(path_to_file):
    os.remove(path_to_file)


def merge_vtus_timestep(base_path, timestep, remove=False, reference_file_path=None):
    """Joins the per-field VTU outputs of one time step ('concentration', 'proliferation', 'growth', 'displacement') with
    the label map into ``merged/all_<step>.vtu`` (data_io.py:606-639)."""
    print("-- Creating joint vtu for timestep %d" % timestep)
    if reference_file_path is None:
        reference_file_path = os.path.join(base_path, "label_map", create_file_name("label_map", 0))
    if os.path.exists(reference_file_path):
        mio_mesh_label = _vtu.read_vtu(reference_file_path)
        for name in ['concentration', 'proliferation', 'growth', 'displacement']:
            path_to_vtu = os.path.join(base_path, name, create_file_name(name, timestep))
            if os.path.exists(path_to_vtu):
                mio_mesh = _vtu.read_vtu(path_to_vtu)
                if name in mio_mesh.point_data.keys():
                    mio_mesh_label.point_data[name] = mio_mesh.point_data[name]
                    if remove:
                        remove_vtu(path_to_vtu)
            else:
                print("   - File '%s' not found" % path_to_vtu)
        path_to_merged = os.path.join(base_path, 'merged', create_file_name("all", timestep))
        print("   - Saving joint file to '%s'" % path_to_merged)
        fu.ensure_dir_exists(path_to_merged)
        _vtu.write_vtu(path_to_merged, mio_mesh_label)
    else:
        print("   - Could not find reference file '%s'... skipping" % reference_file_path)


def merge_VTUs(base_path, delta_t, t_max, remove=False, reference=None):
    """Merges all VTU outputs of a simulation run using `merge_vtus_timestep` (data_io.py:649-654)."""
    for timestep in range(len(np.arange(0, t_max, delta_t)) + 1):
        merge_vtus_timestep(base_path, timestep, remove=remove, reference_file_path=reference)


# ==============================================================================
# MESH IO for parallel processing (data_io.py:663-760)
# ==============================================================================


def save_mesh_hdf5(mesh_in, path_to_file, subdomains=None, boundaries=None):
    """``/mesh`` (+ ``/subdomains`` cell labels, ``/boundaries`` facet labels) in one HDF5 file."""
    hdf = fenics.HDF5File(None, path_to_file, "w")
    hdf.write(mesh_in, "/mesh")
    if subdomains is not None:
        hdf.write(subdomains, "/subdomains")
    if boundaries is not None:
        hdf.write(boundaries, "/boundaries")
    hdf.close()


def read_mesh_hdf5(path_to_file):
    """Returns (mesh, subdomains, boundaries); missing label sets come back as all-zero MeshFunctions."""
    mesh = fenics.Mesh()
    hdf = fenics.HDF5File(None, path_to_file, "r")
    hdf.read(mesh, "/mesh", False)
    subdomains = fenics.MeshFunction("size_t", mesh, mesh.geometry().dim)
    if hdf.has_dataset("subdomains"):
        hdf.read(subdomains, "/subdomains")
    boundaries = fenics.MeshFunction("size_t", mesh, mesh.geometry().dim - 1)
    if hdf.has_dataset("boundaries"):
        hdf.read(boundaries, "/boundaries")
    hdf.close()
    return mesh, subdomains, boundaries


def save_functions_hdf5(function_dict, path_to_file, time_step=None):
    if not function_dict:
        print("No functions provided ... cannot write.")
        return
    hdf = fenics.HDF5File(None, path_to_file, "w")
    for name, function in function_dict.items():
        if time_step is None:
            hdf.write(function, name)
        else:
            hdf.write(function, name, time_step)
    hdf.close()


def read_function_hdf5(name, functionspace, path_to_file):
    if os.path.exists(path_to_file):
        f = fenics.Function(functionspace)
        hdf = fenics.HDF5File(None, path_to_file, "r")
        hdf.read(f, name + "/vector_0")
        hdf.close()
        return f


# ---- small dof / file helpers of the reference that need no image library (data_io.py:132-143, 277-308, 763-800) -------------
def get_value_dimension_from_function(fenics_function):
    """Number of value components (1 for a scalar element, dim for a vector element)."""
    return int(fenics_function.function_space().ncomp)


def get_dof_coordinate_map(functionspace):
    """[n_dofs, gdim] coordinates, one row per dof."""
    return np.asarray(functionspace.tabulate_dof_coordinates()).reshape(-1, functionspace.mesh().geometry().dim())


def get_dofs_by_subspace(functionspace):
    return {i: functionspace.sub(i).dofmap().dofs() for i in range(functionspace.num_sub_spaces())}


def get_dofs_from_coord(dof_coord_map, coord, eps=1e-5):
    """Indices of the dofs within ``eps`` of ``coord`` in every direction (None, with a message, if there is none)."""
    coord = np.asarray(coord, dtype=float)[:dof_coord_map.shape[1]]
    hit = np.nonzero(np.all(np.abs(dof_coord_map - coord[None, :]) < eps, axis=1))[0]
    if len(hit) > 0:
        return hit
    print("Did not find vertex close to (%s) in mesh" % ", ".join(map(str, coord)))
    return None


def save_function_mesh(function, path_to_hdf5_function, labelfunction=None, subdomains=None):
    """``<name>.h5`` (the function, as ``/function``) next to ``<name>_mesh.h5`` (mesh + cell labels)."""
    if not path_to_hdf5_function.endswith(".h5"):
        print("Provide path to '.h5' file")
        return
    path_to_hdf5_mesh = path_to_hdf5_function[:-3] + "_mesh.h5"
    mesh = function.function_space().mesh()
    fu.ensure_dir_exists(path_to_hdf5_mesh)
    if labelfunction is not None:
        from glimslib_b200.simulation_helpers.helper_classes import SubDomains
        sd = SubDomains(mesh)
        sd.setup_subdomains(label_function=labelfunction)
        subdomains = sd.subdomains
    save_mesh_hdf5(mesh, path_to_hdf5_mesh, subdomains=subdomains)
    save_functions_hdf5({"function": function}, path_to_hdf5_function, time_step=None)


def load_function_mesh(path_to_hdf5_function, functionspace="function", degree=1):
    """Inverse of `save_function_mesh`: returns (function, mesh, subdomains, boundaries)."""
    if not path_to_hdf5_function.endswith(".h5"):
        print("Provide path to '.h5' file")
        return None
    path_to_hdf5_mesh = path_to_hdf5_function[:-3] + "_mesh.h5"
    if not os.path.exists(path_to_hdf5_mesh):
        print("Could not find mesh file: '%s'" % path_to_hdf5_mesh)
        return None
    mesh, subdomains, boundaries = read_mesh_hdf5(path_to_hdf5_mesh)
    if functionspace == "function":
        functionspace = fenics.FunctionSpace(mesh, "Lagrange", degree)
    elif functionspace == "vector":
        functionspace = fenics.VectorFunctionSpace(mesh, "Lagrange", degree)
    function = read_function_hdf5("function", functionspace, path_to_hdf5_function)
    return function, mesh, subdomains, boundaries
