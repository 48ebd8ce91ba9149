// Element geometry shared by the assembly kernels (kernels.cu, ccrow.cu): barycentric gradients and volume of a
// P1 simplex, and the exact-integration constants of the mass / cubic reaction terms.
#pragma once
#include "common.h"

namespace {

template <int D> struct Geo { double g[D + 1][D]; double vol; };

__device__ inline void geometry(const double (&X)[3][2], Geo<2>& G) {
    double j00 = X[1][0] - X[0][0], j01 = X[2][0] - X[0][0];
    double j10 = X[1][1] - X[0][1], j11 = X[2][1] - X[0][1];
    double det = j00 * j11 - j01 * j10, id = 1.0 / det;
    G.g[1][0] = j11 * id;  G.g[1][1] = -j01 * id;
    G.g[2][0] = -j10 * id; G.g[2][1] = j00 * id;
    G.g[0][0] = -(G.g[1][0] + G.g[2][0]);
    G.g[0][1] = -(G.g[1][1] + G.g[2][1]);
    G.vol = 0.5 * fabs(det);
}
__device__ inline void geometry(const double (&X)[4][3], Geo<3>& G) {
    double e1[3], e2[3], e3[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { e1[k] = X[1][k] - X[0][k]; e2[k] = X[2][k] - X[0][k]; e3[k] = X[3][k] - X[0][k]; }
    double c1[3] = {e2[1] * e3[2] - e2[2] * e3[1], e2[2] * e3[0] - e2[0] * e3[2], e2[0] * e3[1] - e2[1] * e3[0]};
    double c2[3] = {e3[1] * e1[2] - e3[2] * e1[1], e3[2] * e1[0] - e3[0] * e1[2], e3[0] * e1[1] - e3[1] * e1[0]};
    double c3[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
    double det = e1[0] * c1[0] + e1[1] * c1[1] + e1[2] * c1[2], id = 1.0 / det;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        G.g[1][k] = c1[k] * id; G.g[2][k] = c2[k] * id; G.g[3][k] = c3[k] * id;
        G.g[0][k] = -(G.g[1][k] + G.g[2][k] + G.g[3][k]);
    }
    G.vol = fabs(det) * (1.0 / 6.0);
}

template <int D> struct Consts;
template <> struct Consts<2> { static constexpr double mass = 1.0 / 12.0; static constexpr double kappa = 1.0 / 60.0; };
template <> struct Consts<3> { static constexpr double mass = 1.0 / 20.0; static constexpr double kappa = 1.0 / 120.0; };


}  // namespace
