"""Debug helper: run golden cases on the GPU and save the factors under gpurun_out/ for offline analysis."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from golden_util import draws_from_case, load_case  # noqa: E402

import localmd_b200  # noqa: E402

for name in sys.argv[1:]:
    g, spec, movie = load_case(name)
    d = draws_from_case(g, spec, movie, lazy_sim=True)
    det, tim = {}, {}
    arr = localmd_b200.localmd_decomposition(movie, spec["block_sizes"], spec["frame_range"], draws=d, details=det, timings=tim,
                                             **spec["kwargs"])
    u = arr.u
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "dump_%s.npz" % name), U_data=u.data, U_indices=u.indices,
                        U_indptr=u.indptr, U_shape=np.array(u.shape), R=arr.r, s=arr.s, Vt=arr.v, mean=arr.mean_img, std=arr.var_img,
                        ranks=det["ranks"], sstat=det["sstat"], tstat=det["tstat"], mixing=det["mixing"], v_init=det["v_init"],
                        v_full=det["v_full"], thr=np.array(det["thresholds"]), bg=det["spatial_basis"])
    print(name, "ranks", det["ranks"].tolist(), "k", len(arr.s), {k: round(v, 2) for k, v in tim.items()})
