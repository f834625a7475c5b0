"""Parity metrics between a decomposition of the CUDA path and the CPU oracle (TEST INFRASTRUCTURE: imported only by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs -- never by the product package).

The quantities are the ones BASELINE.json's north_star names: per-block ranks and CSR structure (bit-exact except blocks
with a statistic within EPS of a threshold), singular values (relative error), principal angles of U.R and Vt, and the
relative Frobenius error of the reconstruction Y_hat = U R diag(s) Vt in normalised units."""
import numpy as np

EPS_STAT = 2e-4   # relative band around a threshold inside which a rank decision may legitimately differ


def principal_angles(a, b):
    qa, _ = np.linalg.qr(a)
    qb, _ = np.linalg.qr(b)
    c = np.clip(np.linalg.svd(qa.T @ qb, compute_uv=False), -1, 1)
    return np.arccos(c)


def near_threshold_blocks(sstat, tstat, thr, eps=EPS_STAT):
    return (np.abs(sstat - thr[0]) <= eps * thr[0]).any(axis=1) | (np.abs(tstat - thr[1]) <= eps * thr[1]).any(axis=1)


def recon_rel_err(u_a, r_a, s_a, vt_a, u_b, r_b, s_b, vt_b):
    """|| A - B ||_F / || B ||_F for A = U_a R_a diag(s_a) Vt_a (same for B) WITHOUT forming the d x T matrices:
    <A, B> = sum( (S_a (U_a R_a)^T (U_b R_b) S_b) o (Vt_a Vt_b^T) ), everything in float64."""
    ra = np.asarray(r_a, np.float64) * np.asarray(s_a, np.float64)[None]
    rb = np.asarray(r_b, np.float64) * np.asarray(s_b, np.float64)[None]
    va, vb = np.asarray(vt_a, np.float64), np.asarray(vt_b, np.float64)
    ua, ub = u_a @ ra, u_b @ rb                    # (d, k) dense
    def inner(x, vx, y, vy):
        return float(np.sum((x.T @ y) * (vx @ vy.T)))
    aa, bb, ab = inner(ua, va, ua, va), inner(ub, vb, ub, vb), inner(ua, va, ub, vb)
    return float(np.sqrt(max(aa + bb - 2 * ab, 0.0) / bb))


def parity_report(arr, details, ref, lead_frac=0.05):
    """arr: PMDArray of the CUDA path, details: its `details` dict (ranks, sstat, tstat), ref: OracleResult."""
    out = {}
    ranks_d, ranks_o = np.asarray(details["ranks"]), np.asarray(ref.ranks)
    same = ranks_d == ranks_o
    near = near_threshold_blocks(details["sstat"], details["tstat"], ref.thresholds)
    out["blocks_total"] = int(len(ranks_o))
    out["blocks_rank_equal"] = int(same.sum())
    out["blocks_rank_differ_within_eps"] = int((~same & near).sum())
    out["blocks_rank_differ_outside_eps"] = int((~same & ~near).sum())
    out["eps_stat"] = EPS_STAT
    out["thresholds_rel_err"] = float(np.max(np.abs(np.asarray(details["thresholds"]) / np.asarray(ref.thresholds) - 1)))
    u = arr.u
    if same.all():
        ru = ref.u.copy()
        ru.sort_indices()
        out["csr_indptr_equal"] = bool(np.array_equal(u.indptr, ru.indptr))
        out["csr_indices_equal"] = bool(np.array_equal(u.indices, ru.indices))
    else:
        out["csr_indptr_equal"] = out["csr_indices_equal"] = None   # structure differs where the ranks differ
    k = min(len(arr.s), len(ref.s))
    out["k"] = [int(len(arr.s)), int(len(ref.s))]
    lead = ref.s[:k] > lead_frac * ref.s[0]
    out["s_max_rel_err_lead"] = float(np.max(np.abs(arr.s[:k][lead] / ref.s[:k][lead] - 1)))
    out["s_lead_count"] = int(lead.sum())
    # principal angles of the leading singular subspaces, cut at the widest spectral gap of the lead set
    nl = 1
    if lead.sum() > 1:
        gaps = ref.s[: lead.sum() - 1] / ref.s[1 : lead.sum()]
        nl = int(np.argmax(gaps)) + 1
    out["subspace_dim"] = nl
    ur_d = u @ np.asarray(arr.r[:, :nl], np.float64)
    ur_o = ref.u @ np.asarray(ref.r[:, :nl], np.float64)
    out["angle_UR_max_rad"] = float(principal_angles(ur_d, ur_o).max())
    out["angle_Vt_max_rad"] = float(principal_angles(np.asarray(arr.v[:nl], np.float64).T, np.asarray(ref.vt[:nl], np.float64).T).max())
    out["yhat_rel_fro_err"] = recon_rel_err(u, arr.r, arr.s, arr.v, ref.u, ref.r, ref.s, ref.vt)
    return out


def within_north_star(rep):
    """True when a parity_report meets BASELINE.json's stated tolerances."""
    return bool(rep["blocks_rank_differ_outside_eps"] == 0 and rep["s_max_rel_err_lead"] <= 1e-4
                and rep["angle_UR_max_rad"] <= 1e-3 and rep["angle_Vt_max_rad"] <= 1e-3 and rep["yhat_rel_fro_err"] <= 1e-4
                and rep["csr_indptr_equal"] is not False and rep["csr_indices_equal"] is not False)
