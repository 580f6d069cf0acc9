"""CPU test of the v3 fused step kernel's arithmetic and indexing.

tests/emul/fused3_emul.cpp compiles the kernel's per-lane code
(py-cubed-sphere_b200/csrc/fused3_core.cuh, shared verbatim with the CUDA kernel) with g++
and replays the CTA / warp / lane decomposition and the staged-row ring on the host.  Here
its output for one step is compared with the numpy oracle's divergence + Q update on the
same ghost-filled state (tolerance 1e-13 of max|Q|; the reference bar is 1e-12).
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from golden_common import TUPLES, DT16

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emul", "fused3_emul.cpp")
LIB = os.path.join(HERE, "emul", "libf3emul.so")
CORE = os.path.join(HERE, "..", "py-cubed-sphere_b200", "csrc", "fused3_core.cuh")
JOFF = 12


@pytest.fixture(scope="module")
def emul():
    if (not os.path.exists(LIB)) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(CORE)):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-x", "c++",
                        "-o", LIB, SRC], check=True)
    lib = C.CDLL(LIB)
    dp = C.POINTER(C.c_double)
    lib.f3_emul_step.argtypes = [C.c_int] * 7 + [dp] * 11 + [C.c_double, C.c_int, C.c_double, C.c_double, C.c_double]
    lib.f3_emul_step.restype = C.c_int
    lib.f3_emul_step_block.argtypes = lib.f3_emul_step.argtypes
    lib.f3_emul_step_block.restype = C.c_int
    lib.f3_emul_step_block_circ.argtypes = lib.f3_emul_step.argtypes
    lib.f3_emul_step_block_circ.restype = C.c_int
    lib.f3_emul_step_block_pair.argtypes = lib.f3_emul_step.argtypes
    lib.f3_emul_step_block_pair.restype = C.c_int
    lib.f3_emul_set_variant.argtypes = [C.c_int]
    lib.f3_emul_set_variant.restype = None
    lib.f3_emul_set_block_list.argtypes = [C.POINTER(C.c_int), C.c_int]
    lib.f3_emul_set_block_list.restype = None
    lib.f3_emul_block_grid.argtypes = [C.c_int] * 3
    lib.f3_emul_block_grid.restype = C.c_int
    lib.f3_emul_grid.argtypes = [C.c_int, C.c_int, C.c_int] + [C.POINTER(C.c_int)] * 3
    lib.f3_emul_grid.restype = C.c_int
    lib.f3_emul_ld.argtypes = [C.c_int]
    lib.f3_emul_ld.restype = C.c_int
    return lib


def to_dev(a, ld, single=False):
    """reference layout [i][j][p] -> device layout [p][i][ld] (column j at j + JOFF)."""
    ni, nj, _ = a.shape
    P1 = max(ni, nj)
    npan = 1 if single else 6
    out = np.zeros((npan, P1 if P1 % 2 else P1 + 1, ld))
    out = np.zeros((npan, (min(ni, nj) + 1), ld))
    for p in range(npan):
        out[p, :ni, JOFF:JOFF + nj] = a[:, :, p]
    return out


def ptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def one_step(emul, N, vf, tup, pre_steps, nw=3, depth=3, rows=None, pending=False, separable=False, block_tb=0,
             circ=False, split_sets=None, own_ghosts=None):
    from oracle.grid import LeanGrid
    from oracle import step as ost, wind as owind
    recon, dp, split, et, mt, mf = tup
    g = LeanGrid(N)
    dt = DT16[vf] * 16 / N
    sim = ost.Simulation(g, dt, 5, 2, vf, 1, recon, dp, split, et, mt, mf)
    ost.init_vars_adv(g, sim)
    base_u, base_v = sim.U_pu.ucontra_averaged.copy(), sim.U_pv.vcontra_averaged.copy()
    ost.run(g, sim, pre_steps)
    k = pre_steps + 1
    # the head of adv_time_step (src/advection_timestep.py:28-37)
    ost.ghost_fill_scalar(sim.Q, sim.Q, g, sim)
    if vf >= 2:
        owind.ghost_fill_vector(sim.U_pu, sim.U_pv, sim.U_pc, g, sim)
        owind.time_averaged_velocity(g, sim)
    P = N + 8
    ld = emul.f3_emul_ld(N)
    I = np.s_[4:N + 4, 4:N + 4, :]
    mtc = g.metric_tensor_pc
    Qin = sim.Q.copy()
    corr = 0.0
    if pending:
        # the device holds Q without the previous step's projection term on the interior
        corr = 3.0e-3
        Qin[I] = Qin[I] - mtc[I] * corr
    if separable:
        ws = np.cos(np.pi * ((k - 1) * dt) / 5.0)
        ua, va, mask = base_u, base_v, 2
    else:
        ws = 1.0
        ua, va, mask = sim.U_pu.ucontra_averaged, sim.U_pv.vcontra_averaged, (1 if dp == 2 else 0)
    q = to_dev(Qin, ld)
    qn = np.zeros_like(q)
    rgc = np.zeros_like(mtc)
    rgc[mtc != 0] = 1.0 / mtc[mtc != 0]
    arrs = [to_dev(ua, ld), to_dev(va, ld), to_dev(sim.U_pu.ucontra, ld), to_dev(sim.U_pv.vcontra, ld),
            to_dev(mtc, ld, True), to_dev(rgc, ld, True), to_dev(g.metric_tensor_pu, ld, True),
            to_dev(g.metric_tensor_pv, ld, True)]
    rows = rows or N
    if block_tb:        # v2b decomposition (csrc/fused2b.cu)
        part = np.zeros(emul.f3_emul_block_grid(N, block_tb, rows))
        fn = {False: emul.f3_emul_step_block, True: emul.f3_emul_step_block_circ,
              "pair": emul.f3_emul_step_block_pair}[circ]
        if own_ghosts is not None:
            # one-kernel step (GH = 2): every CTA runs on an input whose ghost cells are NaN except the ones its own
            # prologue fills -- own_ghosts(block) lists them as (i, j) of the block's panel
            for blk in range(len(part)):
                qb = q.copy()
                qb[:, :4, :] = qb[:, N + 4:, :] = np.nan
                qb[:, :, :JOFF + 4] = qb[:, :, JOFF + N + 4:] = np.nan
                pnl = blk % 6
                for i, j in own_ghosts(blk):
                    qb[pnl, i, JOFF + j] = q[pnl, i, JOFF + j]
                lst = np.array([blk], dtype=np.int32)
                emul.f3_emul_set_block_list(lst.ctypes.data_as(C.POINTER(C.c_int)), 1)
                rc = fn(N, recon, split, mask, block_tb, depth, rows, ptr(qb), ptr(qn),
                        *[ptr(a) for a in arrs], ptr(part), corr, 1 if pending else 0, dt / g.dx, dt / g.dy, ws)
                emul.f3_emul_set_block_list(None, 0)
                assert rc == 0
        elif split_sets is None:
            rc = fn(N, recon, split, mask, block_tb, depth, rows, ptr(q), ptr(qn),
                    *[ptr(a) for a in arrs], ptr(part), corr, 1 if pending else 0, dt / g.dx, dt / g.dy, ws)
        else:
            # split step: the interior CTAs run while the ghost cells are still being written -- here: NaN
            inner, outer = [np.ascontiguousarray(x, dtype=np.int32) for x in split_sets]
            qbad = q.copy()
            qbad[:, :4, :] = qbad[:, N + 4:, :] = np.nan
            qbad[:, :, :JOFF + 4] = qbad[:, :, JOFF + N + 4:] = np.nan
            for lst, src in ((inner, qbad), (outer, q)):
                emul.f3_emul_set_block_list(lst.ctypes.data_as(C.POINTER(C.c_int)), len(lst))
                rc = fn(N, recon, split, mask, block_tb, depth, rows, ptr(src), ptr(qn),
                        *[ptr(a) for a in arrs], ptr(part), corr, 1 if pending else 0, dt / g.dx, dt / g.dy, ws)
                emul.f3_emul_set_block_list(None, 0)
                assert rc == 0
    else:               # v3 decomposition (csrc/fused3.cu)
        ns, wc, nch = C.c_int(), C.c_int(), C.c_int()
        npart = emul.f3_emul_grid(N, nw, rows, C.byref(ns), C.byref(wc), C.byref(nch))
        part = np.zeros(npart)
        rc = emul.f3_emul_step(N, recon, split, mask, nw, depth, rows, ptr(q), ptr(qn), *[ptr(a) for a in arrs],
                               ptr(part), corr, 1 if pending else 0, dt / g.dx, dt / g.dy, ws)
    assert rc == 0
    got = np.transpose(qn[:, 4:N + 4, JOFF + 4:JOFF + 4 + N], (1, 2, 0)).copy()
    if mf == 3:                                          # deferred projection (src/discrete_operators.py:98-101)
        a2 = np.sum(mtc[I] * mtc[I])
        got += mtc[I] * (-np.sum(part) / a2)
    # the oracle's step on the same state
    ost.divergence(g, sim)
    want = sim.Q[I] - sim.dt * sim.div[I]
    return got, want


CASES = [
    (16, 1, "default", 2), (20, 1, "default", 0), (50, 3, "default", 3), (130, 1, "default", 1),
    (16, 2, "AVLT-RK2-DG-PR", 3), (50, 2, "AVLT-RK2-DG-PR", 2),
    (16, 1, "PL07-RK1-DG-PR", 2), (50, 3, "PL07-RK1-DG-PR", 1),
    (16, 4, "default", 2),
]


@pytest.mark.parametrize("N,vf,name,pre", CASES)
def test_emulated_kernel_matches_oracle(emul, N, vf, name, pre):
    got, want = one_step(emul, N, vf, TUPLES[name], pre)
    assert np.max(np.abs(got - want)) <= 1e-13 * np.max(np.abs(want))


@pytest.mark.parametrize("tup", [(1, 1, 1, 3, 1, 3), (1, 2, 2, 3, 1, 1), (3, 1, 2, 3, 1, 3), (1, 1, 3, 3, 2, 1)])
def test_emulated_kernel_other_schemes(emul, tup):
    """PPM-0 reconstruction, SP-L04 / SP-PL07 splittings, RK2 masks (tuples outside the named set)."""
    got, want = one_step(emul, 20, 2, tup, 2)
    assert np.max(np.abs(got - want)) <= 1e-13 * np.max(np.abs(want))


@pytest.mark.parametrize("rows,nw,depth", [(7, 3, 1), (16, 3, 2), (9, 4, 3), (50, 2, 4)])
def test_emulated_kernel_chunks_and_shapes(emul, rows, nw, depth):
    """Row chunks shorter than the panel, other CTA widths and rows in flight (depth): same result."""
    got, want = one_step(emul, 50, 3, TUPLES["default"], 2, nw=nw, depth=depth, rows=rows)
    assert np.max(np.abs(got - want)) <= 1e-13 * np.max(np.abs(want))


def test_emulated_kernel_pending_projection(emul):
    """The producer's patch: Q stored without the previous MF-PR term + apply_corr == Q with it."""
    got, want = one_step(emul, 50, 1, TUPLES["default"], 2, pending=True)
    assert np.max(np.abs(got - want)) <= 1e-13 * np.max(np.abs(want))


def test_emulated_kernel_separable_wind(emul):
    """vf = 3 / RK1: wind(0) * cos(pi t / T) scaled inside the kernel."""
    got, want = one_step(emul, 50, 3, TUPLES["default"], 4, separable=True)
    assert np.max(np.abs(got - want)) <= 1e-13 * np.max(np.abs(want))


# ---- the default kernel's decomposition (v2b, csrc/fused2b.cu: one column per thread) ----------
@pytest.mark.parametrize("N,vf,name,pre", CASES)
def test_emulated_v2b_matches_oracle(emul, N, vf, name, pre):
    got, want = one_step(emul, N, vf, TUPLES[name], pre, depth=2, block_tb=32 if N < 100 else 160)
    assert np.max(np.abs(got - want)) <= 1e-13 * np.max(np.abs(want))


@pytest.mark.parametrize("tb,rows,depth,kw", [(32, 7, 1, {}), (64, 16, 2, {"pending": True}),
                                              (160, 50, 3, {"separable": True}), (128, 9, 2, {"pending": True})])
def test_emulated_v2b_shapes_patch_separable(emul, tb, rows, depth, kw):
    got, want = one_step(emul, 50, 3, TUPLES["default"], 3, depth=depth, rows=rows, block_tb=tb, **kw)
    assert np.max(np.abs(got - want)) <= 1e-13 * np.max(np.abs(want))


@pytest.mark.parametrize("tup", [(1, 1, 1, 3, 1, 3), (1, 2, 2, 3, 1, 1), (3, 1, 2, 3, 1, 3), (1, 1, 3, 3, 2, 1)])
def test_emulated_v2b_other_schemes(emul, tup):
    got, want = one_step(emul, 20, 2, tup, 2, depth=2, block_tb=32)
    assert np.max(np.abs(got - want)) <= 1e-13 * np.max(np.abs(want))


# ---- v2b, const-slot march (MINB >= 30): circular six-register windows, phase = row % 6 ------
@pytest.mark.parametrize("N,vf,name,pre", CASES)
def test_emulated_v2b_circular_windows_match_oracle(emul, N, vf, name, pre):
    got, want = one_step(emul, N, vf, TUPLES[name], pre, depth=2, block_tb=32 if N < 100 else 160, circ=True)
    assert np.max(np.abs(got - want)) <= 1e-13 * np.max(np.abs(want))


@pytest.mark.parametrize("tb,rows,kw", [(32, 7, {}), (64, 16, {"pending": True}), (160, 50, {"separable": True}),
                                        (128, 9, {"pending": True}), (160, 6, {}), (64, 12, {})])
def test_emulated_v2b_circular_windows_chunk_lengths(emul, tb, rows, kw):
    """chunks of 13, 22, 56, 15, 12 and 18 marched rows: every exit phase of the six-row group."""
    got, want = one_step(emul, 50, 3, TUPLES["default"], 3, depth=2, rows=rows, block_tb=tb, circ=True, **kw)
    assert np.max(np.abs(got - want)) <= 1e-13 * np.max(np.abs(want))


def test_emulated_v2b_circular_windows_against_shifting(emul):
    """same decomposition, two forms of the y-flux arithmetic (weights / factored): a few ulp apart."""
    a, _ = one_step(emul, 50, 3, TUPLES["default"], 3, depth=2, rows=11, block_tb=64, circ=True)
    b, _ = one_step(emul, 50, 3, TUPLES["default"], 3, depth=2, rows=11, block_tb=64, circ=False)
    assert 0 < np.max(np.abs(a - b)) <= 1e-14 * np.max(np.abs(b))


# ---- march variants of the const-slot kernel (F2B_VAR bits in csrc/fused2b.cu) ---------------------------------
@pytest.mark.parametrize("var", [1, 8, 16, 25])
@pytest.mark.parametrize("tb,rows,kw", [(32, 7, {}), (64, 16, {"pending": True}), (160, 50, {"separable": True}),
                                        (160, 6, {}), (64, 12, {})])
def test_emulated_v2b_march_variants(emul, var, tb, rows, kw):
    """copies of row r+3 issued after barrier B (1), own cell from the register in the y-stencils (8), single
    sqrtg load at the upwind centre (16): same result as the default march (bit for bit unless the mirrored
    stencil of variant 8 reorders the sums) and as the oracle."""
    base, want = one_step(emul, 50, 3, TUPLES["default"], 3, depth=2, rows=rows, block_tb=tb, circ=True, **kw)
    emul.f3_emul_set_variant(var)
    try:
        got, _ = one_step(emul, 50, 3, TUPLES["default"], 3, depth=2, rows=rows, block_tb=tb, circ=True, **kw)
    finally:
        emul.f3_emul_set_variant(0)
    assert np.max(np.abs(got - want)) <= 1e-13 * np.max(np.abs(want))
    if var & 8:
        assert np.max(np.abs(got - base)) <= 1e-14 * np.max(np.abs(base))
    else:
        assert np.array_equal(got, base)


@pytest.mark.parametrize("tup", [(1, 1, 1, 3, 1, 3), (3, 2, 2, 3, 1, 3), (1, 1, 3, 3, 2, 1)])
def test_emulated_v2b_march_variants_other_schemes(emul, tup):
    emul.f3_emul_set_variant(25)
    try:
        got, want = one_step(emul, 20, 2, tup, 2, depth=2, block_tb=32, circ=True)
    finally:
        emul.f3_emul_set_variant(0)
    assert np.max(np.abs(got - want)) <= 1e-13 * np.max(np.abs(want))


# ---- one-kernel step (GH = 2): a CTA reads no ghost cell that its own prologue has not filled -------------------
@pytest.mark.parametrize("N,tb,rows,kw", [(50, 32, 11, {}), (50, 64, 50, {"pending": True}), (20, 32, 7, {}),
                                          (130, 160, 40, {"separable": True})])
def test_emulated_one_kernel_step_ghost_rectangles(emul, N, tb, rows, kw):
    """Every CTA of the uniform grid marched alone on an input whose ghost ring is NaN except for the cells the
    prologue of THAT CTA fills (tests/test_host_cpu.py::_ghost_cells_of_cta mirrors the enumeration of
    csrc/fused2b.cu: ghost_prologue): the step must come out as with the complete ring."""
    from test_host_cpu import _ghost_cells_of_cta
    wmax = tb - 6
    ns = (N + wmax - 1) // wmax
    wcols = (N + ns - 1) // ns
    wcols += wcols & 1

    def own(blk):
        b = blk // 6
        strip, chunk = b % ns, b // ns
        r0 = 4 + chunk * rows
        r1 = min(r0 + rows, N + 4)
        j0 = 4 + strip * wcols
        j1 = min(j0 + wcols, N + 4)
        return _ghost_cells_of_cta(N, r0, r1, j0, j1)

    vf = 3 if kw.get("separable") else 1
    got, want = one_step(emul, N, vf, TUPLES["default"], 2, depth=2, rows=rows, block_tb=tb, circ=True, own_ghosts=own, **kw)
    assert np.all(np.isfinite(got))
    assert np.max(np.abs(got - want)) <= 1e-13 * np.max(np.abs(want))


# ---- v2b, two rows per pair of barriers (MINB >= 50): rings of 4 / 8 slots, windows of 8 registers --
@pytest.mark.parametrize("N,vf,name,pre", CASES)
def test_emulated_v2b_two_row_march_matches_oracle(emul, N, vf, name, pre):
    got, want = one_step(emul, N, vf, TUPLES[name], pre, depth=2, block_tb=32 if N < 100 else 160, circ="pair")
    assert np.max(np.abs(got - want)) <= 1e-13 * np.max(np.abs(want))


@pytest.mark.parametrize("tb,rows,kw", [(32, 7, {}), (64, 16, {"pending": True}), (160, 50, {"separable": True}),
                                        (128, 9, {"pending": True}), (160, 6, {}), (64, 12, {}), (64, 25, {})])
def test_emulated_v2b_two_row_march_chunk_lengths(emul, tb, rows, kw):
    """odd and even numbers of marched rows, every exit phase of the eight-row group, copies issued two steps ahead."""
    got, want = one_step(emul, 50, 3, TUPLES["default"], 3, depth=2, rows=rows, block_tb=tb, circ="pair", **kw)
    assert np.max(np.abs(got - want)) <= 1e-13 * np.max(np.abs(want))


def test_emulated_v2b_two_row_march_same_bits_as_const_slot(emul):
    a, _ = one_step(emul, 50, 3, TUPLES["default"], 3, depth=2, rows=11, block_tb=64, circ="pair")
    b, _ = one_step(emul, 50, 3, TUPLES["default"], 3, depth=2, rows=11, block_tb=64, circ=True)
    assert np.array_equal(a, b)


