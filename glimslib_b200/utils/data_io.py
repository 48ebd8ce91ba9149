"""Mesh / function HDF5 I/O of the drop-in: the subset of ``glimslib/utils/data_io.py`` that sits
directly before and after the hot path -- ``save_mesh_hdf5`` / ``read_mesh_hdf5`` (:663-713, how the MPI examples load
pre-partitioned meshes, README.md:162-183), ``save_functions_hdf5`` / ``read_function_hdf5`` (:715-760).
Image / VTK / meshio conversion stays out of scope (SimpleITK, vtk, meshio are not dependencies here)."""
import os

from glimslib_b200 import fenics_local as fenics


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
