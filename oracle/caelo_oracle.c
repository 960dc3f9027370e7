/* TEST INFRASTRUCTURE ONLY — CPU oracle for the CAE-LO hot path (plain C, scalar).
 *
 * A restatement of the reference's algorithms (SRainGit/CAE-LO) with the arithmetic
 * contract written out, so that a CUDA kernel can be compared bit-for-bit.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library; nothing under caelo_b200/ links or calls it.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC (oracle/Makefile).
 * -ffp-contract=off matters: every fused multiply-add below is an explicit fmaf().
 *
 * Pinning (tests/test_oracle_vs_golden.py, tests/golden/make_golden.py):
 *   respond+select -> DemoData Features/.mat KeyPts (set agreement 1021..1023/1024)
 *                     and bit-identical to the imported reference GetKeyPtsByAE on the same response
 *   nn_match       -> bit-identical to scipy cdist+argmin (the reference's call, Match.py:257-258)
 *   kabsch/ransac  -> the imported reference SolveRT / SolveRelativePose on seeded demo pairs
 *                     (R,t within 1e-4; the reference's own float32 LAPACK noise is the limit)
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------
 * a1  RespondLayer.predict  (SphericalRing.py:405-408, BatchPreprocess.py:110; graph
 *     AE4SphericalRingPC.py:132-133; weights SphericalRingPCRespondLayer.h5)
 *     relu(conv1x1(relu(conv3x3_same(x)+b1))+b2), NHWC, zero padding.
 * Contract R1: acc=b; acc=fmaf(x,w,acc) over (ky,kx,ci) ascending; relu; second layer
 *     acc=b2; acc=fmaf(h[ci],w2[ci][co],acc) over ci ascending; relu.  All float32.
 * ------------------------------------------------------------------------------------ */
void oracle_respond(const float *x, int B, int H, int W, const float *w1 /*3,3,3,32*/,
                    const float *b1, const float *w2 /*32,8*/, const float *b2, float *out)
{
    for (int b = 0; b < B; ++b) {
        const float *xb = x + (size_t)b * H * W * 3;
        float *ob = out + (size_t)b * H * W * 8;
        for (int r = 0; r < H; ++r)
            for (int c = 0; c < W; ++c) {
                float h[32];
                for (int co = 0; co < 32; ++co) h[co] = b1[co];
                for (int ky = 0; ky < 3; ++ky)
                    for (int kx = 0; kx < 3; ++kx) {
                        int rr = r + ky - 1, cc = c + kx - 1;
                        int inside = rr >= 0 && rr < H && cc >= 0 && cc < W;
                        for (int ci = 0; ci < 3; ++ci) {
                            float v = inside ? xb[((size_t)rr * W + cc) * 3 + ci] : 0.0f;
                            const float *w = w1 + ((ky * 3 + kx) * 3 + ci) * 32;
                            for (int co = 0; co < 32; ++co) h[co] = fmaf(v, w[co], h[co]);
                        }
                    }
                for (int co = 0; co < 32; ++co) h[co] = h[co] > 0.0f ? h[co] : 0.0f;
                float *o = ob + ((size_t)r * W + c) * 8;
                for (int c2 = 0; c2 < 8; ++c2) {
                    float acc = b2[c2];
                    for (int ci = 0; ci < 32; ++ci) acc = fmaf(h[ci], w2[ci * 8 + c2], acc);
                    o[c2] = acc > 0.0f ? acc : 0.0f;
                }
            }
    }
}

/* ------------------------------------------------------------------------------------
 * a2  GetKeyPtsByAE  (SphericalRing.py:113-291; restated per SURVEY Appendix D.1)
 * Contract S1 (float32, no FMA): d=nb-ctr; sq=d*d;
 *     n=sqrtf(((sq0+sq1)+(sq2+sq3))+((sq4+sq5)+(sq6+sq7)))   (numpy's 8-lane pairwise tree)
 *     score = min over occupied neighbours in the 5x5 window (centre excluded).
 *     range norm = sqrtf of the SEQUENTIAL float32 sum of squares over all ring channels.
 *     order: ascending by (score, r*W+c); output = candidates[-maxk-1:-1]  (quirk 1).
 * counter_kind: 0 = int8, 1 = int32.  score_out (H*W floats, may be NULL) receives the
 * masked score map (0 where rejected) for diagnostics.
 * Returns the number of keypoints written (<= maxk); kpix is (row, col) int64.
 * ------------------------------------------------------------------------------------ */
typedef struct { uint32_t score_bits; uint32_t idx; } cand_t;

static int cand_cmp(const void *a, const void *b)
{
    const cand_t *x = (const cand_t *)a, *y = (const cand_t *)b;
    if (x->score_bits != y->score_bits) return x->score_bits < y->score_bits ? -1 : 1;
    if (x->idx != y->idx) return x->idx < y->idx ? -1 : 1;
    return 0;
}

static inline int occ_at(const void *counter, int kind, int cW, int r, int c)
{
    if (kind == 0) return ((const int8_t *)counter)[(size_t)r * cW + c] > 0;
    return ((const int32_t *)counter)[(size_t)r * cW + c] > 0;
}

int oracle_select(const float *resp, int H, int W, const float *ring, int ringC, int ringH,
                  int ringW, const void *counter, int counter_kind, int cH, int cW, int maxk,
                  float *kpts, int64_t *kpix, float *score_out)
{
    (void)ringH; (void)cH;
    cand_t *cand = (cand_t *)malloc(sizeof(cand_t) * (size_t)H * W);
    int ncand = 0;
    if (score_out) memset(score_out, 0, sizeof(float) * (size_t)H * W);
    /* quirk 2: rows [8,H-8), cols >= 8 and not in [H-8, H) (the row bound reused for columns) */
    const int edge = 8;
    for (int r = 2; r < H - 2; ++r)
        for (int c = 2; c < W - 2; ++c) {
            int self_ok = occ_at(counter, counter_kind, cW, r, c) && r >= edge && r < H - edge &&
                          c >= edge && !(c >= H - edge && c < H);
            if (!self_ok) continue;
            const float *ctr = resp + ((size_t)r * W + c) * 8;
            float best = INFINITY;
            int count = 0;
            for (int dr = -2; dr <= 2; ++dr)
                for (int dc = -2; dc <= 2; ++dc) {
                    if (dr == 0 && dc == 0) continue;
                    if (!occ_at(counter, counter_kind, cW, r + dr, c + dc)) continue;
                    const float *nb = resp + ((size_t)(r + dr) * W + (c + dc)) * 8;
                    float sq[8];
                    for (int k = 0; k < 8; ++k) {
                        float d = nb[k] - ctr[k];
                        sq[k] = d * d;
                    }
                    float s = ((sq[0] + sq[1]) + (sq[2] + sq[3])) + ((sq[4] + sq[5]) + (sq[6] + sq[7]));
                    float n = sqrtf(s);
                    if (n < best) best = n;
                    ++count;
                }
            if (count < 5) continue;
            /* MinDiffMap_ > 0.2 is evaluated in float64 on a float32 value (SphericalRing.py:173,199) */
            if (!((double)best > 0.2)) continue;
            /* final row/col window (SphericalRing.py:210-213) */
            if (!(r >= edge && r < H - edge && c >= edge && c < W - edge)) continue;
            const float *p = ring + ((size_t)r * ringW + c) * ringC;
            float s = 0.0f;
            for (int k = 0; k < ringC; ++k) s = s + p[k] * p[k];
            if (!(sqrtf(s) >= 10.0f)) continue;
            if (score_out) score_out[(size_t)r * W + c] = best;
            uint32_t bits;
            memcpy(&bits, &best, 4);
            cand[ncand].score_bits = bits;
            cand[ncand].idx = (uint32_t)(r * W + c);
            ++ncand;
        }
    qsort(cand, (size_t)ncand, sizeof(cand_t), cand_cmp);
    /* candidates[-maxk-1:-1] */
    int hi = ncand - 1;
    int lo = ncand - maxk - 1;
    if (lo < 0) lo = 0;
    int n = hi - lo;
    if (n < 0) n = 0;
    for (int i = 0; i < n; ++i) {
        uint32_t idx = cand[lo + i].idx;
        int r = (int)(idx / (uint32_t)W), c = (int)(idx % (uint32_t)W);
        const float *p = ring + ((size_t)r * ringW + c) * ringC;
        kpts[i * 3 + 0] = p[0];
        kpts[i * 3 + 1] = p[1];
        kpts[i * 3 + 2] = p[2];
        kpix[i * 2 + 0] = r;
        kpix[i * 2 + 1] = c;
    }
    free(cand);
    return n;
}

/* ------------------------------------------------------------------------------------
 * a4  cdist(Codes0, Codes1,'euclidean') + argmin(axis=0)   (Match.py:257-258)
 * Contract M1 (float64, no FMA): s=0; for k: d=a[k]-b[k]; s=s+d*d; dist=sqrt(s);
 *     argmin over i, ties -> lowest i.  Verified bit-identical to scipy 1.18 cdist.
 * ------------------------------------------------------------------------------------ */
void oracle_nn_match(const float *c0, int N, const float *c1, int M, int D, int64_t *idx,
                     double *dist_out /* may be NULL */)
{
    for (int j = 0; j < M; ++j) {
        double best = INFINITY;
        int bi = 0;
        const float *b = c1 + (size_t)j * D;
        for (int i = 0; i < N; ++i) {
            const float *a = c0 + (size_t)i * D;
            double s = 0.0;
            for (int k = 0; k < D; ++k) {
                double d = (double)a[k] - (double)b[k];
                s = s + d * d;
            }
            double dist = sqrt(s);
            if (dist < best) { best = dist; bi = i; }
        }
        idx[j] = bi;
        if (dist_out) dist_out[j] = best;
    }
}

/* ------------------------------------------------------------------------------------
 * a5  SolveRT (Match.py:138-158) — Kabsch with the reference's reflection quirk.
 * Contract K1 (float64, no FMA; +,-,*,/,sqrt only so gcc and nvcc agree bit-for-bit):
 *   sums run over the n selected points in 32 interleaved lanes (lane l takes points
 *   l, l+32, ...) each sequential, then the lanes are combined by the xor-butterfly
 *   16,8,4,2,1 (what a warp __shfl_xor reduction computes);
 *   m0,m1 = sums/n;  H[a][b] = sum (p1[a]-m1[a])*(p0[b]-m0[b]);
 *   V = eigenvectors of S=H^T H by cyclic Jacobi, columns sorted by descending eigenvalue;
 *   u_i = Gram-Schmidt(H v_i); if lambda_3 <= 1e-14*lambda_1 the third pair is completed
 *   right-handed (the reference's result is LAPACK rounding noise there);
 *   Q = sum v_i u_i^T (= Vh^T U^T);  det(Q)<0 -> rows scaled by diag(1,1,-1) (quirk 4);
 *   T = m0 - Q m1;  R,T rounded to float32 once.
 * sel: idx!=NULL -> point i is p[idx[i]], i<n;  else mask!=NULL -> points with mask[i]!=0
 *      among the first n;  else all n points.
 * ------------------------------------------------------------------------------------ */
static double lane_tree(const double *lane)
{
    double t[32];
    for (int i = 0; i < 32; ++i) t[i] = lane[i];
    for (int k = 16; k >= 1; k >>= 1)
        for (int i = 0; i < k; ++i) t[i] = t[i] + t[i + k];
    return t[0];
}

static void jacobi3(double S[3][3], double V[3][3])
{
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 12; ++sweep) {
        double off = fabs(S[0][1]) + fabs(S[0][2]) + fabs(S[1][2]);
        if (off == 0.0) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                double apq = S[p][q];
                if (apq == 0.0) continue;
                double theta = (S[q][q] - S[p][p]) / (2.0 * apq);
                double t = 1.0 / (fabs(theta) + sqrt(theta * theta + 1.0));
                if (theta < 0.0) t = -t;
                double c = 1.0 / sqrt(t * t + 1.0);
                double s = t * c;
                /* S <- J^T S J */
                for (int k = 0; k < 3; ++k) {
                    double skp = S[k][p], skq = S[k][q];
                    S[k][p] = c * skp - s * skq;
                    S[k][q] = s * skp + c * skq;
                }
                for (int k = 0; k < 3; ++k) {
                    double spk = S[p][k], sqk = S[q][k];
                    S[p][k] = c * spk - s * sqk;
                    S[q][k] = s * spk + c * sqk;
                }
                for (int k = 0; k < 3; ++k) {
                    double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - s * vkq;
                    V[k][q] = s * vkp + c * vkq;
                }
            }
    }
}

static void cross3(const double a[3], const double b[3], double o[3])
{
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

static double dot3(const double a[3], const double b[3])
{
    return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2];
}

/* any unit vector orthogonal to unit vector a (deterministic) */
static void ortho3(const double a[3], double o[3])
{
    double ax = fabs(a[0]), ay = fabs(a[1]), az = fabs(a[2]);
    double e[3] = {0.0, 0.0, 0.0};
    if (ax <= ay && ax <= az) e[0] = 1.0; else if (ay <= az) e[1] = 1.0; else e[2] = 1.0;
    cross3(a, e, o);
    double n = sqrt(dot3(o, o));
    o[0] = o[0] / n; o[1] = o[1] / n; o[2] = o[2] / n;
}

void oracle_kabsch_from_H(const double Hin[3][3], const double m0[3], const double m1[3],
                          float R[9], float T[3], int *credible)
{
    double H[3][3], S[3][3], V[3][3];
    memcpy(H, Hin, sizeof(H));
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            S[i][j] = (H[0][i] * H[0][j] + H[1][i] * H[1][j]) + H[2][i] * H[2][j];
    jacobi3(S, V);
    double lam[3] = {S[0][0], S[1][1], S[2][2]};
    int ord[3] = {0, 1, 2};
    /* sort descending (stable, fixed network) */
    if (lam[ord[0]] < lam[ord[1]]) { int t = ord[0]; ord[0] = ord[1]; ord[1] = t; }
    if (lam[ord[1]] < lam[ord[2]]) { int t = ord[1]; ord[1] = ord[2]; ord[2] = t; }
    if (lam[ord[0]] < lam[ord[1]]) { int t = ord[0]; ord[0] = ord[1]; ord[1] = t; }
    double v[3][3], u[3][3]; /* v[i] = i-th right singular vector, u[i] = i-th left */
    for (int i = 0; i < 3; ++i)
        for (int k = 0; k < 3; ++k) v[i][k] = V[k][ord[i]];
    double l1 = lam[ord[0]], l2 = lam[ord[1]], l3 = lam[ord[2]];
    const double tiny = 1e-14;
    if (!(l1 > 0.0)) {
        /* H == 0: identity rotation */
        for (int i = 0; i < 3; ++i)
            for (int k = 0; k < 3; ++k) { v[i][k] = (i == k); u[i][k] = (i == k); }
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
            /* rank-deficient: complete so that det(Q) = +1 */
            double cu[3], cv[3];
            cross3(u[0], u[1], cu);
            cross3(v[0], v[1], cv);
            double sgn = dot3(cv, v[2]) < 0.0 ? -1.0 : 1.0; /* handedness of the v frame */
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
        for (int c = 0; c < 3; ++c) Q[2][c] = -Q[2][c]; /* quirk 4: diag(1,1,-1) * Q */
    }
    for (int a = 0; a < 3; ++a) {
        double t = m0[a] - ((Q[a][0] * m1[0] + Q[a][1] * m1[1]) + Q[a][2] * m1[2]);
        T[a] = (float)t;
        for (int c = 0; c < 3; ++c) R[a * 3 + c] = (float)Q[a][c];
    }
    if (credible) *credible = cred;
}

int oracle_kabsch(const float *p0, const float *p1, const uint8_t *mask, const int32_t *idx,
                  int n, float R[9], float T[3], int *credible)
{
    double lane[12][32];
    memset(lane, 0, sizeof(lane));
    int cnt = 0;
    /* pass 1: sums for the means */
    for (int i = 0; i < n; ++i) {
        int j = idx ? idx[i] : i;
        if (!idx && mask && !mask[i]) continue;
        int l = i & 31;
        for (int a = 0; a < 3; ++a) {
            lane[a][l] = lane[a][l] + (double)p0[j * 3 + a];
            lane[3 + a][l] = lane[3 + a][l] + (double)p1[j * 3 + a];
        }
        ++cnt;
    }
    if (cnt == 0) return -1;
    double m0[3], m1[3];
    for (int a = 0; a < 3; ++a) {
        m0[a] = lane_tree(lane[a]) / (double)cnt;
        m1[a] = lane_tree(lane[3 + a]) / (double)cnt;
    }
    double hl[9][32];
    memset(hl, 0, sizeof(hl));
    for (int i = 0; i < n; ++i) {
        int j = idx ? idx[i] : i;
        if (!idx && mask && !mask[i]) continue;
        int l = i & 31;
        double a1[3], a0[3];
        for (int a = 0; a < 3; ++a) {
            a1[a] = (double)p1[j * 3 + a] - m1[a];
            a0[a] = (double)p0[j * 3 + a] - m0[a];
        }
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) hl[a * 3 + b][l] = hl[a * 3 + b][l] + a1[a] * a0[b];
    }
    double H[3][3];
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) H[a][b] = lane_tree(hl[a * 3 + b]);
    oracle_kabsch_from_H((const double(*)[3])H, m0, m1, R, T, credible);
    return cnt;
}

/* ------------------------------------------------------------------------------------
 * a5  RANSAC4RT hypothesis scoring (Match.py:186-194).
 * Contract D1 (float32): q_a = fmaf(R[a][2],z, fmaf(R[a][1],y, R[a][0]*x)) + T[a];
 *     e_a = p0[a]-q_a;  d = sqrtf((e0*e0+e1*e1)+e2*e2);  inlier = d < thr.
 * ------------------------------------------------------------------------------------ */
static inline int inlier_d1(const float *R, const float *T, const float *p0, const float *p1, float thr)
{
    float e[3];
    for (int a = 0; a < 3; ++a) {
        float t = R[a * 3 + 0] * p1[0];
        t = fmaf(R[a * 3 + 1], p1[1], t);
        t = fmaf(R[a * 3 + 2], p1[2], t);
        float q = t + T[a];
        e[a] = p0[a] - q;
    }
    float d = sqrtf((e[0] * e[0] + e[1] * e[1]) + e[2] * e[2]);
    return d < thr;
}

void oracle_ransac_score(const float *p0, const float *p1, int N, const int32_t *sample_idx,
                         int Tn, float thr, int32_t *counts, float *Rt /* [Tn,12] R then T */)
{
    for (int t = 0; t < Tn; ++t) {
        float R[9], T[3];
        int cred;
        /* the four samples occupy lanes 0..3 of the K1 reduction */
        oracle_kabsch(p0, p1, NULL, sample_idx + 4 * t, 4, R, T, &cred);
        int c = 0;
        for (int i = 0; i < N; ++i) c += inlier_d1(R, T, p0 + 3 * i, p1 + 3 * i, thr);
        counts[t] = c;
        memcpy(Rt + 12 * t, R, sizeof(R));
        memcpy(Rt + 12 * t + 9, T, sizeof(T));
    }
}

/* The sequential accept/stop rule of RANSAC4RT for ONE threshold round (Match.py:181-206),
 * replayed over pre-scored hypotheses.  best_n_in carries curNumInliers across ladder rounds
 * (the reference never resets it).  Returns the trial count consumed; *best_t = index of the
 * accepted hypothesis in this round or -1; *ok = isSuccess. */
int oracle_ransac_replay(const int32_t *counts, int Tn, int N, int best_n_in, int *best_t,
                         int *best_n_out, int *ok)
{
    int least = (int)(0.2 * (double)N);
    if (least > 100) least = 100;
    double succ = 0.25 * (double)N;
    int it = 0, bn = best_n_in, bt = -1, success = 0;
    while (it < Tn && ((it < 100) || (it < 500 && (double)bn < succ))) {
        int n = counts[it];
        ++it;
        if (n < least) continue;
        if (n > bn) { bn = n; bt = it - 1; }
        success = 1;
    }
    *best_t = bt;
    *best_n_out = bn;
    *ok = success;
    return it;
}

void oracle_inlier_mask(const float *p0, const float *p1, int N, const float *Rt, float thr,
                        uint8_t *mask)
{
    for (int i = 0; i < N; ++i) mask[i] = (uint8_t)inlier_d1(Rt, Rt + 9, p0 + 3 * i, p1 + 3 * i, thr);
}

/* ------------------------------------------------------------------------------------
 * f1  ProjectPC2SphericalRing  (SphericalRing.py:72-94)
 * Contract P1: r = sqrtf((x*x + y*y) + z*z) in float32 without contraction (numpy's LA.norm
 *     on a float32 (N,3) slice: elementwise square, add.reduce over 3 in order, sqrt); points
 *     with r == 0 are dropped (:77-80); col = (int)((pi - atan2((double)y,(double)x)) / az_res)
 *     and row = ImgH - (int)(asin((double)(z / r)) / v_res + v_off) with z / r a FLOAT32
 *     quotient (both are np.float32 scalars), everything else double, C truncation (:86-88);
 *     rows outside [0,ImgH) are skipped (:89-90); columns are NOT bounds-checked by the
 *     reference (col == ImgW only for atan2 == -pi, where numpy raises IndexError): such
 *     points are counted in the return value and skipped.  The LAST point in file order owns
 *     the pixel (:91-92); the counter counts every hit (:93).
 * ring [ImgH,ImgW,5] f32 and counter [ImgH,ImgW] i32 must be zeroed by the caller.
 * ------------------------------------------------------------------------------------ */
int oracle_project_ring(const float *pc, int64_t N, int ImgH, int ImgW, double az_res, double v_res,
                        double v_off, float *ring, int32_t *counter)
{
    int bad_cols = 0;
    for (int64_t i = 0; i < N; ++i) {
        const float x = pc[4 * i], y = pc[4 * i + 1], z = pc[4 * i + 2];
        const float xx = x * x, yy = y * y, zz = z * z;
        const float s = (xx + yy) + zz;
        const float r = sqrtf(s);
        if (!(r > 0.0f)) continue;
        const int col = (int)((M_PI - atan2((double)y, (double)x)) / az_res);
        const float q = z / r;
        const double beta = asin((double)q);
        const int row = ImgH - (int)(beta / v_res + v_off);
        if (row < 0 || row >= ImgH) continue;
        if (col < 0 || col >= ImgW) { ++bad_cols; continue; }
        float *px = ring + ((size_t)row * ImgW + col) * 5;
        px[0] = x; px[1] = y; px[2] = z; px[3] = pc[4 * i + 3]; px[4] = r;
        counter[(size_t)row * ImgW + col] += 1;
    }
    return bad_cols;
}

/* ------------------------------------------------------------------------------------
 * f2  Voxelization  (Voxel.py:89-173), the parts BatchVoxelization.py:61 writes out.
 * Contract V1 (float64 on float32 inputs, as numpy 1.18 promotes np.float32 + python float):
 *     drop |x| > VisL or |y| > VisW or |z| > VisH (:89-97); x_ = (double)x + VisL ...;
 *     iBlock = (int)(x_ / 1.28); scale-0 voxel = (int)((x_ - iBlock*1.28) / 0.02) + iBlock*64
 *     (the BLOCK route, :120-139 — not (int)(x_/0.02)); scale-1/2 voxel = (int)(x_ / 0.16),
 *     (int)(x_ / 0.64) (:143-148).  A point whose scale-0 voxel was seen before is skipped
 *     entirely (:134-135).  Order: AllVoxels1/2 in first-seen order; AllVoxels0 (and the
 *     in-block AllVoxels) grouped by block in block-first-seen order, first-seen inside (:161-165).
 * Outputs: vox0/vox1/vox2 int16 rows (capacity N each), local0 int16 rows (AllVoxels),
 *     blocks int16 rows (avlBlocksList), cnt int32 [nblocks+1] (cntVoxelsLength);
 *     counts[4] = {V0, V1, V2, nblocks}.  Returns 0, or -1 if a point indexes outside the
 *     156x156x23 block grid / 64^3 block (the reference raises IndexError there).
 * ------------------------------------------------------------------------------------ */
typedef struct { uint64_t key; int32_t val; } vslot;

static uint32_t vhash(uint64_t k)
{
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return (uint32_t)k;
}

/* returns the slot of key, inserting it with val if new (*isnew = 1) */
static vslot *vfind(vslot *t, uint32_t mask, uint64_t key, int32_t val, int *isnew)
{
    uint32_t s = vhash(key) & mask;
    for (;;) {
        if (t[s].key == key) { *isnew = 0; return t + s; }
        if (t[s].key == ~0ull) { t[s].key = key; t[s].val = val; *isnew = 1; return t + s; }
        s = (s + 1) & mask;
    }
}

int oracle_voxelize(const float *pc, int64_t N, int stride, int16_t *vox0, int16_t *vox1, int16_t *vox2,
                    int16_t *local0, int16_t *blocks, int32_t *cnt, int32_t *counts)
{
    const double BRS = 1.28, VS = 0.02;
    const int nBL = (int)(2 * 100 / BRS), nBW = (int)(2 * 100 / BRS), nBH = (int)(2 * 15 / BRS);
    const double VisL = nBL / 2.0 * BRS, VisW = nBW / 2.0 * BRS, VisH = nBH / 2.0 * BRS;
    const double VS1 = VS * 8, VS2 = VS * 32;
    uint32_t cap = 1024;
    while (cap < 2 * (uint64_t)N + 2) cap <<= 1;
    vslot *t0 = malloc(sizeof(vslot) * cap), *t1 = malloc(sizeof(vslot) * cap), *t2 = malloc(sizeof(vslot) * cap),
          *tb = malloc(sizeof(vslot) * cap);
    int32_t *blk_of = malloc(sizeof(int32_t) * (N + 1));   /* block rank of the i-th scale-0 voxel (first-seen order) */
    int16_t *tmp0 = malloc(sizeof(int16_t) * 3 * (N + 1));
    for (uint32_t i = 0; i < cap; ++i) t0[i].key = t1[i].key = t2[i].key = tb[i].key = ~0ull;
    int n0 = 0, n1 = 0, n2 = 0, nb = 0, rc = 0;
    for (int64_t i = 0; i < N; ++i) {
        const float fx = pc[stride * i], fy = pc[stride * i + 1], fz = pc[stride * i + 2];
        if (fabs((double)fx) > VisL || fabs((double)fy) > VisW || fabs((double)fz) > VisH) continue;
        const double x_ = (double)fx + VisL, y_ = (double)fy + VisW, z_ = (double)fz + VisH;
        const int bx = (int)(x_ / BRS), by = (int)(y_ / BRS), bz = (int)(z_ / BRS);
        if (bx < 0 || bx >= nBL || by < 0 || by >= nBW || bz < 0 || bz >= nBH) { rc = -1; continue; }
        const int vx = (int)((x_ - bx * BRS) / VS), vy = (int)((y_ - by * BRS) / VS), vz = (int)((z_ - bz * BRS) / VS);
        if (vx < 0 || vx >= 64 || vy < 0 || vy >= 64 || vz < 0 || vz >= 64) { rc = -1; continue; }
        int isnew;
        const uint64_t kb = (uint64_t)bx | ((uint64_t)by << 16) | ((uint64_t)bz << 32);
        vslot *sb = vfind(tb, cap - 1, kb, nb, &isnew);
        if (isnew) { blocks[3 * nb] = (int16_t)bx; blocks[3 * nb + 1] = (int16_t)by; blocks[3 * nb + 2] = (int16_t)bz; ++nb; }
        const int gx = vx + bx * 64, gy = vy + by * 64, gz = vz + bz * 64;
        const uint64_t k0 = (uint64_t)gx | ((uint64_t)gy << 16) | ((uint64_t)gz << 32);
        vfind(t0, cap - 1, k0, n0, &isnew);
        if (!isnew) continue;
        tmp0[3 * n0] = (int16_t)gx; tmp0[3 * n0 + 1] = (int16_t)gy; tmp0[3 * n0 + 2] = (int16_t)gz;
        blk_of[n0] = sb->val;
        ++n0;
        const int x1 = (int)(x_ / VS1), y1 = (int)(y_ / VS1), z1 = (int)(z_ / VS1);
        const int x2 = (int)(x_ / VS2), y2 = (int)(y_ / VS2), z2 = (int)(z_ / VS2);
        vfind(t1, cap - 1, (uint64_t)x1 | ((uint64_t)y1 << 16) | ((uint64_t)z1 << 32), n1, &isnew);
        if (isnew) { vox1[3 * n1] = (int16_t)x1; vox1[3 * n1 + 1] = (int16_t)y1; vox1[3 * n1 + 2] = (int16_t)z1; ++n1; }
        vfind(t2, cap - 1, (uint64_t)x2 | ((uint64_t)y2 << 16) | ((uint64_t)z2 << 32), n2, &isnew);
        if (isnew) { vox2[3 * n2] = (int16_t)x2; vox2[3 * n2 + 1] = (int16_t)y2; vox2[3 * n2 + 2] = (int16_t)z2; ++n2; }
    }
    /* group scale-0 voxels by block, stable (counting sort on the block rank) */
    for (int b = 0; b <= nb; ++b) cnt[b] = 0;
    for (int i = 0; i < n0; ++i) cnt[blk_of[i] + 1]++;
    for (int b = 0; b < nb; ++b) cnt[b + 1] += cnt[b];
    int32_t *fill = malloc(sizeof(int32_t) * (nb + 1));
    for (int b = 0; b < nb; ++b) fill[b] = cnt[b];
    for (int i = 0; i < n0; ++i) {
        const int b = blk_of[i], o = fill[b]++;
        for (int c = 0; c < 3; ++c) {
            vox0[3 * o + c] = tmp0[3 * i + c];
            local0[3 * o + c] = (int16_t)(tmp0[3 * i + c] - blocks[3 * b + c] * 64);
        }
    }
    counts[0] = n0; counts[1] = n1; counts[2] = n2; counts[3] = nb;
    free(fill); free(tmp0); free(blk_of); free(t0); free(t1); free(t2); free(tb);
    return rc;
}
