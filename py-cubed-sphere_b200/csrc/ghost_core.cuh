// Lagrange (duo-grid) ghost cells, device functions shared by the stand-alone ghost-fill kernel
// (fused.cu) and the ghost prologue of the step kernel (fused2b.cu).
// Reference: src/interpolation.py:154-314; tables src/lagrange.py:28-163; strips src/halo_data.py:15-185.
#pragma once
#include "pycs_common.cuh"

__device__ __forceinline__ double halo_src(const double* __restrict__ q, const Geo& g, const SideMap& m, int a,
                                           int b) {
  return q[gidx(g, m.nb, m.ci + m.ai * a + m.bi * b, m.cj + m.aj * a + m.bj * b)];
}

// One ghost cell of phase 1 (src/interpolation.py:200-248): side s of panel p, ghost layer gl,
// position k along the edge.  Same operations in the same order as dg_phase1_kernel in halo.cu.
__device__ __forceinline__ double dg_phase1_value(const Geo& g, const HaloMaps& maps, const double* __restrict__ q,
                                                  const int* __restrict__ kminE, const double* __restrict__ wE,
                                                  int order, int p, int s, int gl, int k) {
  const SideMap& m = maps.m[p][s];
  const int ge = (s == SIDE_E || s == SIDE_N) ? gl : PYCS_NG - 1 - gl;
  const int km = kminE[ge * g.P + k];
  const double* w = wE + ((long long)ge * g.P + k) * order;
  double acc = 0.0;
  for (int l = 0; l < order; ++l) {
    double v = (s < 2) ? halo_src(q, g, m, gl, km + l) : halo_src(q, g, m, km + l, gl);
    acc = __dadd_rn(acc, __dmul_rn(v, w[l]));
  }
  return acc;
}

// A corner ghost cell of phase 2 (src/interpolation.py:250-314): E or W side s of panel p, layer
// gl, position k in the ghost range.  Its stencil reads the neighbour's strip, whose ends are
// that neighbour's phase-1 ghosts: they are recomputed here (same arithmetic, same bits), so the
// corner depends on interior cells only.
__device__ __forceinline__ double dg_corner_value(const Geo& g, const HaloMaps& maps, const double* __restrict__ q,
                                                  const int* __restrict__ kminE, const double* __restrict__ wE,
                                                  int order, int p, int s, int gl, int k) {
  const SideMap& m = maps.m[p][s];
  const int ge = (s == SIDE_E) ? gl : PYCS_NG - 1 - gl;
  const int km = kminE[ge * g.P + k];
  const double* w = wE + ((long long)ge * g.P + k) * order;
  double acc = 0.0;
  for (int l = 0; l < order; ++l) {
    const int b_ = km + l;
    const int si = m.ci + m.ai * gl + m.bi * b_, sj = m.cj + m.aj * gl + m.bj * b_;
    const bool ii = si >= g.lo && si < g.hi, jj = sj >= g.lo && sj < g.hi;
    double v;
    if (ii && jj) v = q[gidx(g, m.nb, si, sj)];
    else if (ii) v = dg_phase1_value(g, maps, q, kminE, wE, order, m.nb, sj >= g.hi ? SIDE_N : SIDE_S,
                                     sj >= g.hi ? sj - g.hi : sj, si);
    else v = dg_phase1_value(g, maps, q, kminE, wE, order, m.nb, si >= g.hi ? SIDE_E : SIDE_W,
                             si >= g.hi ? si - g.hi : si, sj);
    acc = __dadd_rn(acc, __dmul_rn(v, w[l]));
  }
  return acc;
}

// Any ghost cell (i, j) of panel p (at least one index outside [lo, hi)), raw fill.
__device__ __forceinline__ double dg_ghost_cell(const Geo& g, const HaloMaps& maps, const double* __restrict__ q,
                                                const int* __restrict__ kminE, const double* __restrict__ wE,
                                                int order, int p, int i, int j) {
  const bool ii = i >= g.lo && i < g.hi, jj = j >= g.lo && j < g.hi;
  if (ii) return dg_phase1_value(g, maps, q, kminE, wE, order, p, j >= g.hi ? SIDE_N : SIDE_S,
                                 j >= g.hi ? j - g.hi : j, i);
  if (jj) return dg_phase1_value(g, maps, q, kminE, wE, order, p, i >= g.hi ? SIDE_E : SIDE_W,
                                 i >= g.hi ? i - g.hi : i, j);
  return dg_corner_value(g, maps, q, kminE, wE, order, p, i >= g.hi ? SIDE_E : SIDE_W, i >= g.hi ? i - g.hi : i, j);
}
