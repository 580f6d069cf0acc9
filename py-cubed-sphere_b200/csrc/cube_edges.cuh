// The 12 cube edges as pairs of panel sides (+ flip of the along-edge index).
// One table drives parabola averaging and ghost parabolas (ET-PL07), flux
// averaging (MF-AF) and the RK2 wind line copies (ET-S72/PL07), all of which
// the reference spells out edge by edge:
//   src/edges_treatment.py:38-76, :120-190, :242-278, :311-347.
// dir 0: the edge is an i = const line of that panel (x-parabola / u wind),
// dir 1: j = const (y-parabola / v wind); side 0 = lo (i0 / j0), 1 = hi.
// When both panels touch the edge with the same side their axes run against
// each other: q_L/q_R swap and fluxes / normal winds change sign.
#pragma once

struct EdgeEnd { int dir, side, panel; };
struct CubeEdge { EdgeEnd a, b; int flip; };

__host__ __device__ inline CubeEdge cube_edge(int e) {
  const CubeEdge t[12] = {
      {{0, 1, 0}, {0, 0, 1}, 0}, {{0, 1, 1}, {0, 0, 2}, 0}, {{0, 1, 2}, {0, 0, 3}, 0},
      {{0, 1, 3}, {0, 0, 0}, 0}, {{1, 0, 4}, {1, 1, 0}, 0}, {{0, 1, 4}, {1, 1, 1}, 0},
      {{1, 1, 4}, {1, 1, 2}, 1}, {{0, 0, 4}, {1, 1, 3}, 1}, {{1, 1, 5}, {1, 0, 0}, 0},
      {{1, 0, 1}, {0, 1, 5}, 1}, {{1, 0, 2}, {1, 0, 5}, 1}, {{1, 0, 3}, {0, 0, 5}, 0},
  };
  return t[e];
}
