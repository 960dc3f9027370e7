// kabsch.cuh — contract K1 (oracle/caelo_oracle.c: oracle_kabsch_from_H): the float64 Kabsch / SolveRT arithmetic shared by
// the RANSAC kernels (pose.cu) and the batched ICP (icp.cu).  Reference: Match.py:138-158 `SolveRT`.
// Every translation unit that includes this MUST be compiled with -fmad=false (plain float64 +,-,*,/ and sqrt).
#pragma once

namespace {

__device__ __forceinline__ double warp_tree(double v)
{
    // xor-butterfly 16,8,4,2,1: lane 0 ends with the fixed tree of contract K1 (lane_tree)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void cross3(const double a[3], const double b[3], double o[3])
{
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

__device__ __forceinline__ double dot3(const double a[3], const double b[3])
{
    return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2];
}

__device__ __forceinline__ void ortho3(const double a[3], double o[3])
{
    double ax = fabs(a[0]), ay = fabs(a[1]), az = fabs(a[2]);
    double e[3] = {0.0, 0.0, 0.0};
    if (ax <= ay && ax <= az) e[0] = 1.0; else if (ay <= az) e[1] = 1.0; else e[2] = 1.0;
    cross3(a, e, o);
    double n = sqrt(dot3(o, o));
    o[0] = o[0] / n; o[1] = o[1] / n; o[2] = o[2] / n;
}

__device__ void jacobi3(double S[3][3], double V[3][3])
{
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 12; ++sweep) {
        double off = fabs(S[0][1]) + fabs(S[0][2]) + fabs(S[1][2]);
        if (off == 0.0) break;
#pragma unroll
        for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int q = p + 1; q < 3; ++q) {
                double apq = S[p][q];
                if (apq == 0.0) continue;
                double theta = (S[q][q] - S[p][p]) / (2.0 * apq);
                double t = 1.0 / (fabs(theta) + sqrt(theta * theta + 1.0));
                if (theta < 0.0) t = -t;
                double c = 1.0 / sqrt(t * t + 1.0);
                double s = t * c;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    double skp = S[k][p], skq = S[k][q];
                    S[k][p] = c * skp - s * skq;
                    S[k][q] = s * skp + c * skq;
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    double spk = S[p][k], sqk = S[q][k];
                    S[p][k] = c * spk - s * sqk;
                    S[q][k] = s * spk + c * sqk;
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - s * vkq;
                    V[k][q] = s * vkp + c * vkq;
                }
            }
    }
}

// contract K1 from H, m0, m1 -> float32 R (row-major), T; returns credible (+1 / -1)
__device__ int kabsch_from_H(const double H[3][3], const double m0[3], const double m1[3],
                             float R[9], float T[3])
{
    double S[3][3], V[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            S[i][j] = (H[0][i] * H[0][j] + H[1][i] * H[1][j]) + H[2][i] * H[2][j];
    jacobi3(S, V);
    double lam[3] = {S[0][0], S[1][1], S[2][2]};
    int o0 = 0, o1 = 1, o2 = 2, tmp;
    if (lam[o0] < lam[o1]) { tmp = o0; o0 = o1; o1 = tmp; }
    if (lam[o1] < lam[o2]) { tmp = o1; o1 = o2; o2 = tmp; }
    if (lam[o0] < lam[o1]) { tmp = o0; o0 = o1; o1 = tmp; }
    const int ord[3] = {o0, o1, o2};
    double v[3][3], u[3][3];
    for (int i = 0; i < 3; ++i)
        for (int k = 0; k < 3; ++k) {
            // V[k][ord[i]] with a runtime column index, selected without local-memory indexing
            int c = ord[i];
            v[i][k] = c == 0 ? V[k][0] : (c == 1 ? V[k][1] : V[k][2]);
        }
    const double l1 = o0 == 0 ? lam[0] : (o0 == 1 ? lam[1] : lam[2]);
    const double l2 = o1 == 0 ? lam[0] : (o1 == 1 ? lam[1] : lam[2]);
    const double l3 = o2 == 0 ? lam[0] : (o2 == 1 ? lam[1] : lam[2]);
    const double tiny = 1e-14;
    if (!(l1 > 0.0)) {
        for (int i = 0; i < 3; ++i)
            for (int k = 0; k < 3; ++k) { v[i][k] = (i == k) ? 1.0 : 0.0; u[i][k] = (i == k) ? 1.0 : 0.0; }
    } else {
        double b[3];
        for (int k = 0; k < 3; ++k) b[k] = (H[k][0] * v[0][0] + H[k][1] * v[0][1]) + H[k][2] * v[0][2];
        double n = sqrt(dot3(b, b));
        for (int k = 0; k < 3; ++k) u[0][k] = b[k] / n;
        if (l2 > tiny * l1) {
            for (int k = 0; k < 3; ++k) b[k] = (H[k][0] * v[1][0] + H[k][1] * v[1][1]) + H[k][2] * v[1][2];
            double p = dot3(b, u[0]);
            for (int k = 0; k < 3; ++k) b[k] = b[k] - p * u[0][k];
            n = sqrt(dot3(b, b));
            for (int k = 0; k < 3; ++k) u[1][k] = b[k] / n;
        } else {
            ortho3(u[0], u[1]);
        }
        if (l3 > tiny * l1 && l2 > tiny * l1) {
            for (int k = 0; k < 3; ++k) b[k] = (H[k][0] * v[2][0] + H[k][1] * v[2][1]) + H[k][2] * v[2][2];
            double p0 = dot3(b, u[0]);
            for (int k = 0; k < 3; ++k) b[k] = b[k] - p0 * u[0][k];
            double p1 = dot3(b, u[1]);
            for (int k = 0; k < 3; ++k) b[k] = b[k] - p1 * u[1][k];
            n = sqrt(dot3(b, b));
            for (int k = 0; k < 3; ++k) u[2][k] = b[k] / n;
        } else {
            double cu[3], cv[3];
            cross3(u[0], u[1], cu);
            cross3(v[0], v[1], cv);
            double sgn = dot3(cv, v[2]) < 0.0 ? -1.0 : 1.0;
            for (int k = 0; k < 3; ++k) u[2][k] = sgn * cu[k];
        }
    }
    double Q[3][3];
    for (int a = 0; a < 3; ++a)
        for (int c = 0; c < 3; ++c)
            Q[a][c] = (v[0][a] * u[0][c] + v[1][a] * u[1][c]) + v[2][a] * u[2][c];
    double det = (Q[0][0] * (Q[1][1] * Q[2][2] - Q[1][2] * Q[2][1]) -
                  Q[0][1] * (Q[1][0] * Q[2][2] - Q[1][2] * Q[2][0])) +
                 Q[0][2] * (Q[1][0] * Q[2][1] - Q[1][1] * Q[2][0]);
    int cred = 1;
    if (det < 0.0) {
        cred = -1;
        for (int c = 0; c < 3; ++c) Q[2][c] = -Q[2][c];  // quirk 4 (Match.py:151-155)
    }
    for (int a = 0; a < 3; ++a) {
        double t = m0[a] - ((Q[a][0] * m1[0] + Q[a][1] * m1[1]) + Q[a][2] * m1[2]);
        T[a] = (float)t;
        for (int c = 0; c < 3; ++c) R[a * 3 + c] = (float)Q[a][c];
    }
    return cred;
}

}  // namespace
