"""The reference's .npz layout (README.md:25-56, demos/official_demo.ipynb cells 8 and 10).

Keys: fov_shape, fov_order, U_data, U_indices, U_indptr, U_shape, U_format, R, s, Vt, mean_img,
noise_var_img.  The notebook stores U_format = type(U) (a pickled class, hence allow_pickle=True on
load); here it is stored as the string "csr_matrix"; files are loaded WITHOUT pickle, and files written by
the reference's notebook load too (U_format is never read back, NpzFile is lazy)."""
import numpy as np
import scipy.sparse

from .pmdarray import PMDArray


def save_npz(path, arr: PMDArray):
    u = arr.u.tocsr()
    np.savez(
        path,
        fov_shape=np.array(arr.shape[1:]),
        fov_order=arr.order,
        U_data=u.data,
        U_indices=u.indices,
        U_indptr=u.indptr,
        U_shape=np.array(u.shape),
        U_format="csr_matrix",
        R=arr.r,
        s=arr.s,
        Vt=arr.v,
        mean_img=arr.mean_img,
        noise_var_img=arr.var_img,
    )


def load_npz(path, device=None) -> PMDArray:
    # allow_pickle=False: the arrays this loader reads are plain numeric / string arrays also in files written by the
    # reference's notebook (only its U_format entry is a pickled class, and NpzFile never touches entries not asked for)
    data = np.load(path, allow_pickle=False)
    u = scipy.sparse.csr_matrix((data["U_data"], data["U_indices"], data["U_indptr"]), shape=tuple(data["U_shape"]))
    v = data["Vt"]
    shape = (v.shape[1], int(data["fov_shape"][0]), int(data["fov_shape"][1]))
    return PMDArray(u, data["R"], data["s"], v, shape, data["fov_order"].item(), data["mean_img"], data["noise_var_img"],
                    device=device)
