"""CPU tests of the host-side logic and of the C-ABI library itself (no compute calls without a GPU):
the shared library loads and exports every symbol include/pmd_sm100.h declares, the host tables
(tiling, weights, task lists, supertiles, Welch tables) agree with the oracle, the .npz layout and the
PMDArray container work, and compute entry points fail loudly without a CUDA device."""
import os
import tempfile

import numpy as np
import pytest
import scipy.sparse as sp
import torch

import oracle.pmd_oracle as O


def test_library_exports_every_declared_symbol():
    import __graft_entry__

    __graft_entry__.build()
    from localmd_b200 import _lib

    names = _lib.declared_symbols()
    assert len(names) >= 17 and "pmd_project_supertile" in names and "pmd_stats_pass" in names
    L = _lib.lib()
    for n in names:
        assert hasattr(L, n), n
    assert L.pmd_abi_version() == 1
    sigs = _lib._parse_header()
    assert sigs["pmd_reconstruct"] == "pppplplpppp"


def test_null_pointer_is_rejected_without_touching_the_gpu():
    from localmd_b200 import _lib

    L = _lib.lib()
    rc = L.pmd_stats_pass(None, 0, 10, 10, 10, None, None, None, None, None)
    assert rc < 0 and b"null pointer" in L.pmd_last_error()
    rc = L.pmd_gram_f64(None, 1, 500, 10, 1, 1, 1, None, None)
    assert rc < 0


def test_host_geometry_matches_oracle():
    from localmd_b200 import decomposition as D

    for n, b in [(512, 20), (512, 32), (1024, 40), (150, 32), (150, 28), (150, 40), (24, 24), (60, 20), (80, 20)]:
        assert D.tile_starts(n, b) == O.tile_starts(n, b)
    for bh, bw in [(20, 20), (12, 20), (32, 32), (10, 14)]:
        np.testing.assert_array_equal(D.pyramid_weights(bh, bw), O.pyramid_weights(bh, bw))
    with pytest.raises(ValueError):
        D.pyramid_weights(21, 20)
    with pytest.raises(ValueError):
        D.check_fov_size((9, 20))
    with pytest.raises(ValueError):
        D.update_block_sizes([9, 20], (100, 100))
    assert D.update_block_sizes([32, 32], (24, 20)) == [24, 20]
    rng = np.random.default_rng(0)
    fr = D.identify_window_chunks(600, 1300, 600, rng)
    assert len(fr) == 600 and fr[0] in (0, 600, 700) and fr == list(range(fr[0], fr[0] + 600))
    assert D.identify_window_chunks(400, 1300, 200, rng, starting_points=[800, 200]) == list(range(200, 400)) + list(range(800, 1000))
    with pytest.raises(ValueError):
        D.identify_window_chunks(2000, 1300, 2000, rng)


def test_task_tables():
    from localmd_b200 import ops

    assert ops.make_tasks([3, 5, 1, 8]).tolist() == [[0, 0], [1, 0], [1, 4], [2, 0], [3, 0], [3, 4]]
    assert [ops.split_groups(k) for k in (1, 4, 5, 6, 7, 9, 13)] == [[1], [4], [3, 2], [3, 3], [4, 3], [3, 3, 3], [4, 3, 3, 3]]
    rng = np.random.default_rng(1)
    for d1, d2, bh, bw in [(512, 512, 20, 20), (61, 83, 16, 16), (150, 150, 32, 28), (20, 20, 20, 20), (64, 40, 20, 12)]:
        rows, cols = O.tile_starts(d1, bh), O.tile_starts(d2, bw)
        nb = len(rows) * len(cols)
        ranks = rng.integers(1, 14, nb)
        col0 = np.concatenate([[0], np.cumsum(ranks)[:-1]])
        st = ops.make_supertiles(rows, cols, bh, bw, ranks, col0)
        assert st["max_h"] * st["max_w"] <= 2048 or st["G"] == 1
        covered = np.zeros(int(ranks.sum()), dtype=int)
        for ti, (r0, c0, rh, rw) in enumerate(st["tiles"]):
            assert r0 + rh <= d1 and c0 + rw <= d2
            for qi0, qj0, col, nc in st["tasks"][st["task_ptr"][ti] : st["task_ptr"][ti + 1]]:
                assert 0 <= qi0 and qi0 + bh <= rh and 0 <= qj0 and qj0 + bw <= rw and 1 <= nc <= 4
                b = np.searchsorted(col0, col, side="right") - 1
                assert (r0 + qi0, c0 + qj0) == (rows[b // len(cols)], cols[b % len(cols)])
                covered[col : col + nc] += 1
        assert np.all(covered == 1)  # every kept component is projected exactly once


def test_welch_tables_identity():
    """Both kernel formulations of the Welch estimate (folded tables, v1; FFT + real split, v2) equal scipy's welch."""
    from localmd_b200._tables import welch_fft_reference, welch_fft_tables, welch_from_tables_reference, welch_tables

    tc, ts = welch_tables()
    assert tc.shape == (128, 64) and ts.shape == (128, 64) and tc.dtype == np.float32
    assert welch_fft_tables().shape == (772,) and welch_fft_tables().dtype == np.float32
    rng = np.random.default_rng(0)
    for n in (1024, 544, 256, 300):
        x = (200 + 3 * rng.standard_normal((16, n))).astype(np.float32)
        np.testing.assert_allclose(welch_from_tables_reference(x), O.welch_noise_estimate(x), rtol=1e-6)
        np.testing.assert_allclose(welch_fft_reference(x), O.welch_noise_estimate(x), rtol=1e-6)


def test_welch_tc_tables():
    """The decimation-in-frequency form the tensor-core stats kernel evaluates equals scipy's welch, and the operand image
    decodes (TF32 part + bf16 pair part, swizzled K-major layout) back to the two DFT matrices."""
    from localmd_b200._tables import welch_dif_matrices, welch_dif_reference, welch_tc_tables

    rng = np.random.default_rng(1)
    for n in (1024, 544, 256, 300):
        x = (200 + 3 * rng.standard_normal((16, n)) + 0.02 * np.arange(n)).astype(np.float32)
        np.testing.assert_allclose(welch_dif_reference(x), O.welch_noise_estimate(x), rtol=1e-6)
    tab = welch_tc_tables()
    assert tab.dtype == np.uint8 and tab.shape == (131072 + 512,)
    words = tab[:131072].view(np.uint32)
    w = tab[131072:].view(np.float32)
    np.testing.assert_allclose(w, 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(128) / 256), rtol=1e-6, atol=1e-8)
    for par, mat in enumerate(welch_dif_matrices()):
        for n_, k_ in ((0, 0), (5, 37), (63, 127), (17, 64), (40, 95)):
            ka, kk = k_ // 32, k_ % 32
            off = ka * 8192 + (n_ >> 3) * 1024 + (n_ & 7) * 128 + (((kk // 4) ^ (n_ & 7)) << 4) + 4 * (kk % 4)
            hi = words[par * 8192 + off // 4 : par * 8192 + off // 4 + 1].view(np.float32)[0]
            pair = int(words[16384 + par * 8192 + off // 4])
            lo = np.array([(pair & 0xFFFF) << 16], np.uint32).view(np.float32)[0]
            hi_b = np.array([pair & 0xFFFF0000], np.uint32).view(np.float32)[0]
            assert (np.float32(hi).view(np.uint32) & 0x1FFF) == 0                      # exact in TF32
            assert abs(float(hi) + float(lo) - mat[n_, k_]) <= 2.0 ** -17 * max(abs(mat[n_, k_]), 1e-3)
            assert abs(float(hi_b) - float(hi)) <= 2.0 ** -8 * abs(float(hi)) + 1e-30


def test_npz_layout_and_pmdarray_container():
    import localmd_b200

    rng = np.random.default_rng(0)
    d1, d2, T, R, k = 6, 5, 20, 7, 4
    u = sp.random(d1 * d2, R, density=0.4, random_state=1, format="csr")
    arr = localmd_b200.PMDArray(u.tocoo(), rng.standard_normal((R, k)).astype(np.float32), np.array([4, 3, 2, 1], np.float32),
                                rng.standard_normal((k, T)).astype(np.float32), (T, d1, d2), "F", np.ones((d1, d2), np.float32),
                                np.ones((d1, d2), np.float32))
    assert arr.shape == (T, d1, d2) and arr.ndim == 3 and arr.dtype == np.float32 and sp.isspmatrix_csr(arr.u)
    np.testing.assert_array_equal(arr.row_indices, np.arange(30).reshape((6, 5), order="F"))
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "x.npz")
        localmd_b200.save_npz(path, arr)
        data = np.load(path, allow_pickle=True)
        assert sorted(data.files) == sorted(["fov_shape", "fov_order", "U_data", "U_indices", "U_indptr", "U_shape", "U_format",
                                             "R", "s", "Vt", "mean_img", "noise_var_img"])
        assert tuple(data["fov_shape"]) == (d1, d2) and data["fov_order"].item() == "F"
        back = localmd_b200.load_npz(path)
        assert back.shape == arr.shape and (back.u != arr.u).nnz == 0
        np.testing.assert_array_equal(back.v, arr.v)
    for bad in [None, (None, 1, 2), (1, 2, 3, 4)]:
        with pytest.raises(ValueError):
            arr[bad]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_cuda():
    import localmd_b200

    movie = np.zeros((300, 20, 20), np.float32)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        localmd_b200.localmd_decomposition(movie, [16, 16], 100)
    arr = localmd_b200.PMDArray(sp.eye(400, 3, format="csr"), np.eye(3, dtype=np.float32), np.ones(3, np.float32),
                                np.ones((3, 300), np.float32), (300, 20, 20), "F", np.zeros((20, 20), np.float32),
                                np.ones((20, 20), np.float32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        arr[0]


def test_lazy_data_loader_contract():
    import localmd_b200

    data = np.arange(10 * 4 * 3, dtype=np.float32).reshape(10, 4, 3)

    class Arr(localmd_b200.lazy_data_loader):
        dtype = "float32"
        shape = data.shape

        def _compute_at_indices(self, idx):
            return data[idx]

    a = Arr()
    np.testing.assert_array_equal(a[[1, 3]], data[[1, 3]])
    np.testing.assert_array_equal(a[2], data[2])
    np.testing.assert_array_equal(a[2:5, 1], data[2:5, 1])
    np.testing.assert_array_equal(a[np.array([0, 9]), :, 2], data[[0, 9], :, 2])
    with pytest.raises(IndexError):
        a[0, 0, 0, 0]
    with pytest.raises(IndexError):
        a[0:11]
    with pytest.raises(IndexError):
        a["x"]


def test_overlap_pairs_host():
    from localmd_b200 import ops as host_ops

    starts = np.array([(k, j) for k in O.tile_starts(52, 20) for j in O.tile_starts(33, 20)], dtype=np.int32)
    pairs = host_ops.overlap_pairs(starts, 20, 20)
    hit = (np.abs(starts[:, None, 0] - starts[None, :, 0]) < 20) & (np.abs(starts[:, None, 1] - starts[None, :, 1]) < 20)
    b1, b2 = np.nonzero(hit)
    assert np.array_equal(pairs, np.stack([b1, b2], axis=1))
    shuffled = starts[::-1].copy()
    pairs2 = host_ops.overlap_pairs(shuffled, 20, 20)  # arbitrary order -> generic path
    hit2 = (np.abs(shuffled[:, None, 0] - shuffled[None, :, 0]) < 20) & (np.abs(shuffled[:, None, 1] - shuffled[None, :, 1]) < 20)
    assert np.array_equal(pairs2, np.stack(np.nonzero(hit2), axis=1))


def test_strip_tables():
    """make_strips: every (block, component) and every (pixel, background component) is owned by exactly one task;
    the tasks of one warp slot never overlap in rows; extra passes only stream the rows they need."""
    from localmd_b200 import ops as host_ops

    rng = np.random.default_rng(5)
    for (d1, d2, bh, bw, n_bg, hi) in [(112, 95, 20, 20, 15, 14), (61, 83, 10, 10, 1, 4), (70, 96, 32, 32, 9, 30), (40, 40, 40, 40, 3, 5)]:
        rs, cs = O.tile_starts(d1, bh), O.tile_starts(d2, bw)
        ranks = rng.integers(1, hi + 1, len(rs) * len(cs))
        col0 = np.concatenate([[0], np.cumsum(ranks)[:-1]])
        st = host_ops.make_strips(rs, cs, bh, bw, d1, d2, ranks, col0, n_bg)
        ref = host_ops.make_strips_py(rs, cs, bh, bw, d1, d2, ranks, col0, n_bg)  # native builder == Python specification
        for key in ("items", "slot_ptr", "tasks", "local8", "local4"):
            assert np.array_equal(st[key], ref[key]), key
        assert (st["upack_floats"], st["n_parts"], st["max_rw"], st["n_items"]) == (
            ref["upack_floats"], ref["n_parts"], ref["max_rw"], ref["n_items"])
        items, slot_ptr, tasks = st["items"], st["slot_ptr"], st["tasks"]
        assert st["max_rw"] <= host_ops.PS_MAX_RW
        seen_cols = np.zeros(int(ranks.sum()), int)
        bg_cover = np.zeros((n_bg, d1, d2), int)
        for (c0, rw, sp, n_rows, part, row0, _, _) in items:
            for w in range(host_ops.PS_WARPS):
                end = -1
                for ti in range(slot_ptr[sp + w], slot_ptr[sp + w + 1]):
                    by, bx, h, wd, col, nc, ncp, urow, ulo, uhi, kind, _ = tasks[ti]
                    assert by >= end and by >= row0 and by + h <= row0 + n_rows and bx + wd <= rw and 1 <= nc <= ncp <= 8
                    end = by + h
                    if kind == 0:
                        seen_cols[col : col + nc] += 1
                        assert (h, wd) == (bh, bw) and urow == bw * ncp
                    else:
                        bg_cover[col : col + nc, by : by + h, c0 + bx : c0 + bx + wd] += 1
        assert np.all(seen_cols == 1) and np.all(bg_cover == 1)


def test_strip_tables_tc():
    """make_strips_tc: a NumPy emulation of the tensor-core projection kernel driven ONLY by the host tables (slot
    accumulators, coefficient rows, drain events) reproduces U^T Y, so every (block, component) and (pixel, background
    component) is owned exactly once, slots never hold two live tasks, and every task is drained at its last row."""
    from localmd_b200 import ops as host_ops

    rng = np.random.default_rng(11)
    cases = [(112, 96, 20, 20, 15, 14, None), (61, 84, 10, 10, 1, 4, None), (70, 96, 32, 32, 9, 30, None), (40, 40, 40, 40, 3, 5, None),
             (60, 80, 20, 20, 0, 50, None), (64, 64, 16, 16, 5, 3, 2)]
    for (d1, d2, bh, bw, n_bg, hi, G) in cases:
        rs, cs = O.tile_starts(d1, bh), O.tile_starts(d2, bw)
        ranks = rng.integers(1, hi + 1, len(rs) * len(cs))
        col0 = np.concatenate([[0], np.cumsum(ranks)[:-1]])
        n_local = int(ranks.sum())
        st = host_ops.make_strips_tc(rs, cs, bh, bw, d1, d2, ranks, col0, n_bg, G=G)
        assert st is not None
        uv = rng.standard_normal((n_local, bh * bw))
        bg = rng.standard_normal((max(n_bg, 1), d1 * d2))
        T = 3
        y = rng.standard_normal((T, d1, d2))
        # reference
        ref = np.zeros((n_local + n_bg, T))
        blocks = [(a, c) for a in rs for c in cs]
        for b, (i0, j0) in enumerate(blocks):
            for c in range(ranks[b]):
                ref[col0[b] + c] = np.einsum("tij,ij->t", y[:, i0 : i0 + bh, j0 : j0 + bw], uv[col0[b] + c].reshape(bh, bw))
        for k in range(n_bg):
            ref[n_local + k] = y.reshape(T, -1) @ bg[k]
        # emulation
        items, slot_ptr, tasks, events = st["items"], st["slot_ptr"], st["tasks"], st["events"]
        z = np.full((n_local, T), np.nan)
        zbg = np.zeros((st["n_parts"], max(n_bg, 1), T))
        chunk = 0
        for (c0, w8, row0, n_rows, b0, nkc, ev0, n_ev, part, sp0, _, _) in items:
            W = 8 * w8
            assert c0 % 4 == 0 and 0 <= c0 and c0 + W <= d2 and W <= 128 and nkc == (w8 + 3) // 4 and b0 == chunk
            chunk += n_rows * nkc
            acc = np.zeros((host_ops.TC_SLOTS * 4, T))
            e = ev0
            for row in range(row0, row0 + n_rows):
                B = np.zeros((host_ops.TC_SLOTS * 4, W))
                for s in range(host_ops.TC_SLOTS):
                    live = [ti for ti in range(slot_ptr[sp0 + s], slot_ptr[sp0 + s + 1]) if tasks[ti][0] <= row < tasks[ti][0] + tasks[ti][2]]
                    assert len(live) <= 1
                    for ti in live:
                        by, bx, h, wd, col, nc, kind, _ = tasks[ti]
                        assert 1 <= nc <= 4 and bx >= 0 and bx + wd <= W
                        for c in range(nc):
                            if kind == 0:
                                B[4 * s + c, bx : bx + wd] = uv[col + c].reshape(bh, bw)[row - by]
                            else:
                                B[4 * s + c, bx : bx + wd] = bg[col + c].reshape(d1, d2)[row, c0 + bx : c0 + bx + wd]
                acc += B @ y[:, row, c0 : c0 + W].T
                while e < ev0 + n_ev and events[e][0] == row:
                    _, s, col, ncw = events[e]
                    nc, kind = ncw & 0xFF, ncw >> 8
                    if kind == 0:
                        assert np.all(np.isnan(z[col : col + nc]))  # written exactly once
                        z[col : col + nc] = acc[4 * s : 4 * s + nc]
                    else:
                        zbg[part, col : col + nc] += acc[4 * s : 4 * s + nc]
                    acc[4 * s : 4 * s + 4] = 0
                    e += 1
            assert e == ev0 + n_ev and np.all(acc == 0)  # everything drained
        assert chunk == st["chunks"]
        np.testing.assert_allclose(z, ref[:n_local], atol=1e-9)
        if n_bg:
            np.testing.assert_allclose(zbg.sum(0)[:n_bg], ref[n_local:], atol=1e-9)


def test_append_components_and_csr_relabelling_cpu():
    """Device-agnostic host-side tensor logic, run on CPU tensors: the ragged per-block append used by the windowed block
    fits, and SparseU.csr(row_ids): the relabelled CSR is a permutation of the row segments of the physical CSR and must
    equal scipy's canonical CSR of the same matrix with rows numbered in Fortran order."""
    import scipy.sparse as sp
    import torch

    from localmd_b200.decomposition import SparseU, _append_components

    final = torch.zeros((3, 4, 8))
    counter = torch.tensor([0, 2, 5])
    comps = torch.arange(3 * 4 * 8, dtype=torch.float32).reshape(3, 4, 8) + 1
    n_new = torch.tensor([3, 0, 2])
    _append_components(final, counter, comps, n_new)
    assert torch.equal(final[0, :, :3], comps[0, :, :3]) and torch.all(final[0, :, 3:] == 0)
    assert torch.all(final[1] == 0)
    assert torch.equal(final[2, :, 5:7], comps[2, :, :2]) and torch.all(final[2, :, :5] == 0) and torch.all(final[2, :, 7:] == 0)

    rng = np.random.default_rng(8)
    d1, d2, bh, bw, K = 14, 12, 8, 6, 2
    rows, cols = O.tile_starts(d1, bh), O.tile_starts(d2, bw)
    starts = np.array([(a, c) for a in rows for c in cols], dtype=np.int32)
    ranks = rng.integers(1, 4, len(starts)).astype(np.int64)
    n_local = int(ranks.sum())
    uv = rng.standard_normal((n_local, bh * bw))
    uv[rng.uniform(size=uv.shape) < 0.1] = 0.0  # exact zeros are dropped like scipy does
    bg = rng.standard_normal((K, d1 * d2)).astype(np.float32)
    su = SparseU(starts, torch.from_numpy(starts), bh, bw, d1, d2, ranks, torch.from_numpy(ranks.astype(np.int32)),
                 torch.from_numpy(uv), torch.from_numpy(uv.astype(np.float32)), torch.from_numpy(bg))
    dense = np.zeros((d1 * d2, n_local + K))
    col = 0
    qi, qj = np.divmod(np.arange(bh * bw), bw)
    for b, (i0, j0) in enumerate(starts):
        pix = (i0 + qi) * d2 + j0 + qj
        for c in range(ranks[b]):
            dense[pix, col] = uv[col]
            col += 1
    dense[:, n_local:] = bg.T.astype(np.float64)
    for order in ("C", "F"):
        row_ids = np.arange(d1 * d2).reshape((d1, d2), order=order).reshape(-1)  # physical pixel -> row id
        ref = np.zeros_like(dense)
        ref[row_ids] = dense
        ref = sp.csr_matrix(ref)
        ref.sort_indices()
        ip, ix, v = su.csr(torch.from_numpy(row_ids))
        np.testing.assert_array_equal(ip.numpy(), ref.indptr)
        np.testing.assert_array_equal(ix.numpy(), ref.indices)
        np.testing.assert_array_equal(v.numpy(), ref.data)
    ip, ix, v = su.csr()
    refp = sp.csr_matrix(dense)
    refp.sort_indices()
    np.testing.assert_array_equal(ip.numpy(), refp.indptr)
    np.testing.assert_array_equal(ix.numpy(), refp.indices)


def test_utu_host_tables_match_numpy_formulation():
    """pmd_utu_host_tables (native host routine, no device work) against the NumPy statement of the same bookkeeping:
    row offsets of every block-pair tile and the CSR row pointer of U_loc^T U_loc."""
    from localmd_b200 import ops

    rng = np.random.default_rng(0)
    rows, cols = list(range(0, 100, 10)), list(range(0, 90, 10))
    starts = np.stack(np.meshgrid(rows, cols, indexing="ij"), -1).reshape(-1, 2)
    ranks = rng.integers(0, 6, len(starts)).astype(np.int64)
    pairs, rowoff, rowptr = ops.utu_host_tables(starts, 20, 20, ranks)
    b1, b2 = pairs[:, 0].astype(np.int64), pairs[:, 1].astype(np.int64)
    r2 = ranks[b2]
    ex = np.cumsum(r2) - r2
    first = np.searchsorted(b1, np.arange(len(ranks)))
    np.testing.assert_array_equal(rowoff, ex - ex[first[b1]])
    width = np.bincount(b1, weights=r2, minlength=len(ranks)).astype(np.int64)
    np.testing.assert_array_equal(rowptr, np.concatenate([[0], np.cumsum(np.repeat(width, ranks))]))


def test_sym_splits_fill_whole_waves():
    """Split count of pmd_sym_product_f64: units = upper tiles x splits should fill whole waves of 148 SMs."""
    from localmd_b200 import ops

    s = ops.sym_splits(1650, 20000)            # 91 upper tiles: 13 splits -> 1183 units = 7.99 waves
    assert s == 13 and 91 * s <= 8 * 148
    assert ops.sym_splits(100, 500) == 1       # one tile, short inner dimension: no split
    for n, k in [(330, 4096), (3860, 30000), (129, 100000)]:
        s = ops.sym_splits(n, k)
        assert 1 <= s <= 64 and (s == 1 or k // s >= 256)


def test_block_orth_fits_matches_kernel_footprint():
    """Shared-memory footprint of pmd_block_orth (csrc/orth.cu: block_orth_smem) as restated in ops.block_orth_fits."""
    from localmd_b200 import ops

    assert ops.block_orth_fits(400, 50) and ops.block_orth_fits(100, 60)
    assert 2 * ((56 * 56 + 3 * 56) * 8 + 400 * 52 * 4 + 1024) <= 228 * 1024      # two 400 x 50 matrices per SM
    assert not ops.block_orth_fits(1024, 50) and not ops.block_orth_fits(400, 65)


def test_jacobi_block_fold_enumerates_upper_triangle_once():
    """Index arithmetic of jacobi_eigh_kernel (csrc/dense_small.cu): item m of the folded rectangle -> 2 x 2 block (k, l) with
    k <= l; every block of the upper triangle of the (N/2) x (N/2) block grid must be produced exactly once."""
    for half in range(1, 57):
        fold_cols = half if half & 1 else half + 1
        n_blocks = half * (half + 1) // 2
        seen = set()
        for m in range(n_blocks):
            a, c = divmod(m, fold_cols)
            if a + c < half:
                k, l = a, a + c
            elif half & 1:
                k, l = half - a, c
            else:
                k = half - 1 - a
                l = k + (c - (half - a))
            assert 0 <= k <= l < half and (k, l) not in seen, (half, m, k, l)
            seen.add((k, l))
        assert len(seen) == n_blocks
