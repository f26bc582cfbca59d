/* ilqg_backpass_split.cuh -- K2s: backward pass with G lanes per problem for SMALL state dimensions and SMALL batches.
 *
 * Why: with one lane per problem (k_backpass) a warp issues ~1250 instructions per time step, a step is one long dependent
 * chain, and a warp executes the box-QP iterations of its slowest lane.  That is the right trade when the batch fills the
 * GPU (no redundant work at all), but with a few thousand problems per GPU there is less than one warp per scheduler and the
 * pass costs 500 steps x the latency of one warp-step whatever the batch.  Here G = 4 lanes share a problem (8 problems per
 * warp): the matrix work of a step is split by OUTPUT ELEMENT -- lane c owns column c of fx, hence Qx[c], row c of Qxu,
 * column c of Vxx*fx, of the gains L, of Quu*L, Vx[c] and the diagonal entry (c,c) of Qxx / Vxx; the off-diagonal entries of
 * the symmetric matrices are dealt round-robin -- while the short vectors the box QP needs (Qu, Quu) and the QP itself are
 * evaluated redundantly by every lane (identical inputs, identical results, uniform control flow inside a group).  The few
 * values another lane needs (columns of Vxx*fx, L, Quu*L, rows of Qxu, the new value function) travel through shared
 * memory, four group-level __syncwarp per step.  A warp-step is ~950 instructions (one lane per problem: ~1210), and the
 * box-QP iterations of a warp diverge over 4 (default) or 8 problems instead of 32.
 *
 * Measured on B200 (car, 4096 problems, T = 500, 50 passes; ms per backward pass): one lane per problem 2.38 (0.95 in the
 * first passes, 2.3-2.6 once a few lanes per warp iterate the box QP); this kernel with 8 problems per warp 1.88, with 4
 * problems per warp 1.82, with 2 problems per warp 2.8 (more warps than the SMs hold at once).  Both kernels take the same
 * 0.95 ms in the first passes: a step is one dependent chain of ~3700 cycles (about half of it the box QP: Cholesky, inverse
 * and the Armijo test are serial divisions and square roots) that fewer instructions do not shorten, so the gain is the
 * divergence.  The barriers name the four lanes of a group only: groups of a warp drift apart freely; a variant with
 * warp-wide barriers (all groups forced back together four times per step) measured 7 % slower.  From ~8000 problems on
 * the GPU the lane-per-problem kernel is as fast or faster (16 384: 2.4 against 3.9 ms), hence the threshold in ilqg_host.c.
 *
 * Parity: every output element is still ONE serial sum in the reference's order (matMult.c:3-72, back_pass.c:80-241), so the
 * results are bit-identical to k_backpass and to the reference.  The lane-owned products are evaluated densely (a lane's
 * column is not known at compile time, so the structural-zero masks of k_backpass cannot prune them): exactly what the
 * reference does, i.e. the one documented deviation of k_backpass (non-finite x structural zero) does not occur here.
 *
 * Scope: FULL_DDP = 0 and 1, no state-dependent input limits (HAS_HX), lane-per-problem problems (not COOP); everything else
 * keeps using k_backpass / k_backpass_warp.  Reference: back_pass.c:38-257, boxQP.c:39-238, iLQG.c:261-303. */
#pragma once

namespace ilqg {

constexpr int SP_BLOCK = 32;   /* one warp per block: 32 / G problems; blocks spread over all SMs even for small batches */

template <class P> __host__ __device__ constexpr bool split_supported() { return !P::HAS_HX && !use_coop<P>(); }

/* per-problem shared-memory workspace, as offsets (in doubles) into one flat array; every section starts 16-byte aligned and
   the total is = 2 (mod 16) doubles, so the eight groups of a warp start four banks apart: their same-offset accesses never
   conflict */
template <class P> struct SplitWS {
    static constexpr int NX = P::NX, NU = P::NU;
    static constexpr int ev(int n) { return (n + 1) & ~1; }
    static constexpr int OFF_D = 0;                                  /* Dense<P>: constants once, varying entries every step */
    static constexpr int OFF_BA = ev(P::DENSE_SIZE);                 /* Vxx*fx, column c at [c*NX] */
    static constexpr int OFF_LK = OFF_BA + ev(NX * NX);              /* gains L, column c at [c*NU] */
    static constexpr int OFF_BL = OFF_LK + ev(NU * NX);              /* Quu*L, column c at [c*NU] */
    static constexpr int OFF_QT = OFF_BL + ev(NU * NX);              /* Qxu transposed: row i of Qxu at [i*NU] */
    static constexpr int OFF_V = OFF_QT + ev(NU * NX);               /* new Vx | Vxx (packed upper triangle) */
    static constexpr int OFF_V2 = OFF_V + ev(NX + P::NQXX);          /* FULL_DDP: time-varying second-order entries of the step */
    static constexpr int RAW = OFF_V2 + ev(P::NV2);
    static constexpr int SIZE = RAW + (18 - RAW % 16);
};

template <class P, int G> __host__ __device__ constexpr size_t split_smem_bytes() { return sizeof(double) * SplitWS<P>::SIZE * (SP_BLOCK / G); }

/* N doubles from shared memory; 16-byte loads when the caller can promise alignment (even offset, even N) */
template <int N, bool AL> __device__ __forceinline__ void lds_vec(const double *p, double *out)
{
    if (AL && (N % 2 == 0)) {
        const double2 *p2 = reinterpret_cast<const double2 *>(p);
#pragma unroll
        for (int i = 0; i < N / 2; i++) {
            const double2 v = p2[i];
            out[2 * i] = v.x;
            out[2 * i + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < N; i++) out[i] = p[i];
    }
}
template <int N, bool AL> __device__ __forceinline__ void sts_vec(double *p, const double *in)
{
    if (AL && (N % 2 == 0)) {
        double2 *p2 = reinterpret_cast<double2 *>(p);
#pragma unroll
        for (int i = 0; i < N / 2; i++) p2[i] = make_double2(in[2 * i], in[2 * i + 1]);
    } else {
#pragma unroll
        for (int i = 0; i < N; i++) p[i] = in[i];
    }
}

template <class P, bool FULL, bool PP, int G>
__global__ void __launch_bounds__(SP_BLOCK) k_backpass_split(Work w, Opts o, ParamBlock<P> pb, int iter)
{
    static_assert(G == 2 || G == 4 || G == 8, "lanes per problem");
    static_assert(split_supported<P>(), "k_backpass_split: problem class not supported");
    constexpr int NX = P::NX, NU = P::NU, NQXX = P::NQXX, NQUU = P::NQUU, NQXU = P::NQXU;
    constexpr int CPL = (NX + G - 1) / G;             /* columns per lane */
    constexpr int NOFF = NQXX - NX;                   /* off-diagonal entries of a symmetric NX x NX matrix */
    constexpr int ROUNDS = (NOFF + G - 1) / G;        /* off-diagonal entries per lane */
    constexpr int R1 = (P::NV1 + G - 1) / G;          /* time-varying derivative entries each lane fetches */
    constexpr int NV2U = FULL ? P::NV2_USED : 0, R2 = (NV2U + G - 1) / G;   /* ... and second-order entries (FULL_DDP) */
    constexpr bool ALX = (NX % 2 == 0), ALU = (NU % 2 == 0);
    using WS = SplitWS<P>;
    static_assert(sizeof(Dense<P>) == sizeof(double) * P::DENSE_SIZE, "Dense layout must match the generator's table");
    extern __shared__ double2 sp_smem2[];
    const int g = threadIdx.x & (G - 1), grp = threadIdx.x / G;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));
    /* o.bp_ppw problems per warp (at most 32 / G): the first groups of more warps when the batch is small */
    const int ppw = o.bp_ppw < SP_BLOCK / G ? o.bp_ppw : SP_BLOCK / G;
    if (grp >= ppw) return;
    const int b = blockIdx.x * ppw + grp;
    if (b >= w.B) return;
    if (w.status[b] != ST_RUNNING) return;
    ILQG_PARAMS(PP, b)
    double *ws = reinterpret_cast<double *>(sp_smem2) + (size_t)grp * WS::SIZE;
    Dense<P> &D = *reinterpret_cast<Dense<P> *>(ws + WS::OFF_D);
    double *Dd = ws + WS::OFF_D;
    if (w.new_deriv[b]) {
        if (w.deriv_fail[b]) { /* "Calculating derivatives failed": break (iLQG.c:248-251) */
            if (g == 0) finish(w, b, iter, w.bp_done[b] ? 1 : 0);
            return;
        }
        __syncwarp(gmask);
        if (g == 0) {
            w.new_deriv[b] = 0;
            w.n_dv[b] += 1;
        }
    }
    const size_t Bp = w.Bp;
    const int T = w.T;
    const int cur = w.cur[b];
    double lambda = w.lambda[b], dlambda = w.dlambda[b];
    if (g == 0) P::consts(pv, D);

    /* what this lane owns (loop-invariant): columns, off-diagonal entries, derivative entries it fetches */
    int col[CPL];
    bool colv[CPL];
#pragma unroll
    for (int t = 0; t < CPL; t++) {
        colv[t] = (g + G * t) < NX;
        col[t] = colv[t] ? (g + G * t) : 0;
    }
    int orow[ROUNDS > 0 ? ROUNDS : 1], ocol[ROUNDS > 0 ? ROUNDS : 1];
    bool offv[ROUNDS > 0 ? ROUNDS : 1];
#pragma unroll
    for (int t = 0; t < ROUNDS; t++) {
        const int e = g + G * t;   /* e-th off-diagonal entry in packed order: (0,1) (0,2) (1,2) (0,3) ... */
        offv[t] = e < NOFF;
        int c = 1, r = 0;
        if (offv[t]) {
            while ((c * (c + 1)) / 2 <= e) c++;
            r = e - (c * (c - 1)) / 2;
        }
        orow[t] = r;
        ocol[t] = c;
    }
    int dst[R1];
    bool dstv[R1];
#pragma unroll
    for (int t = 0; t < R1; t++) {
        const int j = g + G * t;
        dstv[t] = j < P::NV1;
        dst[t] = dstv[t] ? P::v1_dst(j) : 0;
    }

    double Vx[NX], Vxx[NQXX];
    double dV0 = 0.0, dV1 = 0.0, g_sum = 0.0;
    int n_bp = w.n_bp[b];
    bool done = false;
    while (!done) {
        n_bp++;
#pragma unroll
        for (int i = 0; i < NX; i++) Vx[i] = w.FD[(size_t)i * Bp + b];
#pragma unroll
        for (int i = 0; i < NQXX; i++) Vxx[i] = w.FD[(size_t)(NX + i) * Bp + b];
        dV0 = 0.0;
        dV1 = 0.0;
        g_sum = 0.0;
        double lk[NU];
#pragma unroll
        for (int i = 0; i < NU; i++) lk[i] = 0.0;
        bool failed = false;
        /* the entries of step k-1 are requested while step k is computed (registers pf / un_n) */
        double pf[R1], pf2[R2 > 0 ? R2 : 1], un_n[NU];
        {
            const double *v1 = w.V1 + (size_t)(T - 1) * P::NV1 * Bp + b;
#pragma unroll
            for (int t = 0; t < R1; t++) pf[t] = dstv[t] ? v1[(size_t)(g + G * t) * Bp] : 0.0;
            const double *v2g = w.V2 + (size_t)(T - 1) * P::NV2 * Bp + b;
#pragma unroll
            for (int t = 0; t < R2; t++) pf2[t] = (g + G * t) < NV2U ? v2g[(size_t)(g + G * t) * Bp] : 0.0;
            const double *un = w.XU[cur] + ((size_t)(T - 1) * Bp + b) * Rec<P>::RXU + NX;
#pragma unroll
            for (int i = 0; i < NU; i++) un_n[i] = un[i];
        }
        __syncwarp(gmask);

        for (int k = T - 1; k >= 0; k--) {
#pragma unroll
            for (int t = 0; t < R1; t++)
                if (dstv[t]) Dd[dst[t]] = pf[t];
#pragma unroll
            for (int t = 0; t < R2; t++)
                if ((g + G * t) < NV2U) ws[WS::OFF_V2 + g + G * t] = pf2[t];
            double un[NU];
#pragma unroll
            for (int i = 0; i < NU; i++) un[i] = un_n[i];
            if (k > 0) {
                const double *v1 = w.V1 + (size_t)(k - 1) * P::NV1 * Bp + b;
#pragma unroll
                for (int t = 0; t < R1; t++)
                    if (dstv[t]) pf[t] = v1[(size_t)(g + G * t) * Bp];
                const double *v2g = w.V2 + (size_t)(k - 1) * P::NV2 * Bp + b;
#pragma unroll
                for (int t = 0; t < R2; t++)
                    if ((g + G * t) < NV2U) pf2[t] = v2g[(size_t)(g + G * t) * Bp];
                const double *unp = w.XU[cur] + ((size_t)(k - 1) * Bp + b) * Rec<P>::RXU + NX;
#pragma unroll
                for (int i = 0; i < NU; i++) un_n[i] = unp[i];
            }
            __syncwarp(gmask);

            /* ---- every lane: Qu, Vxx*fu, Quu (back_pass.c:80-131; the operands the box QP waits for) ---- */
            double fu[NX * NU], Qu[NU], Quu[NQUU], bcu[NX * NU];
#pragma unroll
            for (int i = 0; i < NX * NU; i++) fu[i] = mnz<typename P::Mask_fu>(i) ? D.fu[i] : 0.0;
#pragma unroll
            for (int i = 0; i < NU; i++) Qu[i] = D.cu[i];
            add_mul_vec<NX, NU, typename P::Mask_fu>(Qu, Vx, fu);
#pragma unroll
            for (int j = 0; j < NU; j++)
#pragma unroll
                for (int r = 0; r < NX; r++) {
                    double acc = 0.0;
#pragma unroll
                    for (int s = 0; s < NX; s++)
                        if (mnz<typename P::Mask_fu>(s + j * NX)) acc += Vxx[symtri(r, s)] * fu[s + j * NX];
                    bcu[r + j * NX] = acc;
                }
#pragma unroll
            for (int c = 0; c < NU; c++)
#pragma unroll
                for (int r = 0; r <= c; r++) {
                    double acc = 0.0;
#pragma unroll
                    for (int s = 0; s < NX; s++)
                        if (mnz<typename P::Mask_fu>(s + r * NX)) acc += fu[s + r * NX] * bcu[s + c * NX];
                    if (r != c) {
#pragma unroll
                        for (int s = 0; s < NX; s++)
                            if (mnz<typename P::Mask_fu>(s + c * NX)) acc += fu[s + c * NX] * bcu[s + r * NX];
                        acc *= 0.5;
                    }
                    Quu[utri(r, c)] = D.cuu[utri(r, c)] + acc;
                }

            /* ---- own columns: Qx[c], column c of Vxx*fx, row c of Qxu, Qxx(c,c) ---- */
            double Qx_c[CPL], Qxu_c[CPL][NU], Qxx_d[CPL];
#pragma unroll
            for (int t = 0; t < CPL; t++) {
                const int c = col[t];
                double a[NX], bac[NX];
                lds_vec<NX, ALX>(D.fx + c * NX, a);
                double q = D.cx[c];
#pragma unroll
                for (int r = 0; r < NX; r++) q += Vx[r] * a[r];
                Qx_c[t] = q;
#pragma unroll
                for (int r = 0; r < NX; r++) {
                    double acc = 0.0;
#pragma unroll
                    for (int s = 0; s < NX; s++) acc += Vxx[symtri(r, s)] * a[s];
                    bac[r] = acc;
                }
#pragma unroll
                for (int j = 0; j < NU; j++) {
                    double acc = 0.0;
#pragma unroll
                    for (int s = 0; s < NX; s++) acc += a[s] * bcu[s + j * NX];
                    Qxu_c[t][j] = D.cxu[c + j * NX] + acc;
                }
                {
                    double acc = 0.0;
#pragma unroll
                    for (int s = 0; s < NX; s++) acc += a[s] * bac[s];
                    Qxx_d[t] = D.cxx[(c * (c + 1)) / 2 + c] + acc;
                }
                if (colv[t]) sts_vec<NX, ALX>(ws + WS::OFF_BA + c * NX, bac);
            }
            __syncwarp(gmask);
            /* ---- own off-diagonal entries of Qxx (matMult.c:14-46: both halves in ONE accumulator, then * 0.5) ---- */
            double Qxx_o[ROUNDS > 0 ? ROUNDS : 1];
#pragma unroll
            for (int t = 0; t < ROUNDS; t++) {
                const int r = orow[t], c = ocol[t];
                double ar[NX], ac[NX], br[NX], bc[NX];
                lds_vec<NX, ALX>(D.fx + r * NX, ar);
                lds_vec<NX, ALX>(D.fx + c * NX, ac);
                lds_vec<NX, ALX>(ws + WS::OFF_BA + r * NX, br);
                lds_vec<NX, ALX>(ws + WS::OFF_BA + c * NX, bc);
                double acc = 0.0;
#pragma unroll
                for (int s = 0; s < NX; s++) acc += ar[s] * bc[s];
#pragma unroll
                for (int s = 0; s < NX; s++) acc += ac[s] * br[s];
                acc *= 0.5;
                Qxx_o[t] = D.cxx[(c * (c + 1)) / 2 + r] + acc;
            }

            if (FULL) {
                /* ---- FULL_DDP tensor terms (back_pass.c:95-131): every lane evaluates the generated term sums of ALL entries into
                        zeroed arrays (entries without terms stay literal zeros and fold away), then adds those of the entries it
                        owns: Q = (first-order part) + sum, the order of the lane-per-problem kernel ---- */
                double v2[P::NV2], tQxu[P::NQXU], tQuu[NQUU], tQxx[NQXX];
#pragma unroll
                for (int i = 0; i < P::NV2; i++) v2[i] = i < NV2U ? ws[WS::OFF_V2 + i] : 0.0;
#pragma unroll
                for (int i = 0; i < P::NQXU; i++) tQxu[i] = 0.0;
#pragma unroll
                for (int i = 0; i < NQUU; i++) tQuu[i] = 0.0;
#pragma unroll
                for (int i = 0; i < NQXX; i++) tQxx[i] = 0.0;
                P::add2_Qxu(Vx, v2, pv, tQxu);
                P::add2_Quu(Vx, v2, pv, tQuu);
                P::add2_Qxx(Vx, v2, pv, tQxx);
#pragma unroll
                for (int i = 0; i < NQUU; i++) Quu[i] += tQuu[i];
#pragma unroll
                for (int t = 0; t < CPL; t++) {
#pragma unroll
                    for (int j = 0; j < NU; j++) {
                        double d = 0.0;
#pragma unroll
                        for (int c = 0; c < NX; c++) d = (col[t] == c) ? tQxu[c + j * NX] : d;
                        Qxu_c[t][j] += d;
                    }
                    double d = 0.0;
#pragma unroll
                    for (int c = 0; c < NX; c++) d = (col[t] == c) ? tQxx[utri(c, c)] : d;
                    Qxx_d[t] += d;
                }
#pragma unroll
                for (int t = 0; t < ROUNDS; t++) {
                    double d = 0.0;
#pragma unroll
                    for (int c = 1; c < NX; c++)
#pragma unroll
                        for (int r = 0; r < c; r++) d = (ocol[t] == c && orow[t] == r) ? tQxx[utri(r, c)] : d;
                    Qxx_o[t] += d;
                }
            }
            /* ---- regularisation (back_pass.c:134-159) ---- */
            double QuuF[NQUU], Qxu_reg[CPL][NU];
#pragma unroll
            for (int i = 0; i < NQUU; i++) QuuF[i] = Quu[i];
#pragma unroll
            for (int t = 0; t < CPL; t++)
#pragma unroll
                for (int j = 0; j < NU; j++) Qxu_reg[t][j] = Qxu_c[t][j];
            if (o.regType == 2) {
                double fuf[NX * NU];   /* all of fu, structural zeros included: the reference's index pattern reads across columns */
#pragma unroll
                for (int i = 0; i < NX * NU; i++) fuf[i] = D.fu[i];
#pragma unroll
                for (int j = 0; j < NU; j++)
#pragma unroll
                    for (int i = 0; i <= j; i++) {
                        double acc = 0.0;
#pragma unroll
                        for (int c = 0; c < NU; c++) acc += fuf[symtri(c, i)] * fuf[symtri(c, j)];
                        QuuF[utri(i, j)] += acc * lambda;
                    }
#pragma unroll
                for (int t = 0; t < CPL; t++) {
                    double a[NX];
                    lds_vec<NX, ALX>(D.fx + col[t] * NX, a);
#pragma unroll
                    for (int j = 0; j < NU; j++) {
                        double acc = 0.0;
#pragma unroll
                        for (int c = 0; c < NX; c++) acc += a[c] * fuf[(c + j * NU) < NX * NU ? (c + j * NU) : 0];
                        Qxu_reg[t][j] += acc * lambda;
                    }
                }
            }
            if (o.regType == 1) {
#pragma unroll
                for (int i = 0; i < NU; i++) QuuF[utri(i, i)] += lambda;
            }
            /* ---- box QP, warm-started from step k+1, in every lane (back_pass.c:163-171) ---- */
            int clamped[NU], n_free;
            double invH[NQUU];
            int qp;
            {
                double lo[NU], hi[NU];
#pragma unroll
                for (int i = 0; i < NU; i++) {
                    lo[i] = D.lower[i];
                    hi[i] = D.upper[i];
                }
                qp = box_qp<NU>(QuuF, Qu, lo, hi, lk, clamped, invH, n_free);
            }
            if (w.tr_clamp && g == 0) {
                int code = (qp & 0xff) << 16;
#pragma unroll
                for (int i = 0; i < NU; i++) code |= clamped[i] << (2 * i);
                w.tr_clamp[(size_t)k * Bp + b] = code;
            }
            if (qp < 1) { /* the failed QP's last iterate stays in t->l, as in the reference's in-place boxQP */
                if (g == 0) {
                    double *rec = w.Ll[cur] + ((size_t)k * Bp + b) * Rec<P>::RLS;
#pragma unroll
                    for (int i = 0; i < NU; i++) rec[i] = lk[i];
                }
                failed = true;
                break;
            }
            /* ---- gains, own columns (back_pass.c:173-201), and the control-law record of step k ---- */
            double Lk_c[CPL][NU];
            {
                double *recL = w.LL[cur] + ((size_t)k * Bp + b) * Rec<P>::RLM;
#pragma unroll
                for (int t = 0; t < CPL; t++) {
#pragma unroll
                    for (int i = 0; i < NU; i++) {
                        double acc = 0.0;
#pragma unroll
                        for (int j = 0; j < NU; j++) {
                            const double v = acc - invH[symtri(i, j)] * Qxu_reg[t][j];
                            acc = (clamped[i] || clamped[j]) ? acc : v;
                        }
                        Lk_c[t][i] = acc;
                    }
                    if (colv[t]) {
                        sts_vec<NU, ALU>(ws + WS::OFF_LK + col[t] * NU, Lk_c[t]);
                        sts_vec<NU, ALU>(ws + WS::OFF_QT + col[t] * NU, Qxu_c[t]);
                        if (ALU && (Rec<P>::RLM % 2 == 0)) {
                            double2 *r2 = reinterpret_cast<double2 *>(recL + col[t] * NU);
#pragma unroll
                            for (int i = 0; i < NU / 2; i++) r2[i] = make_double2(Lk_c[t][2 * i], Lk_c[t][2 * i + 1]);
                        } else {
#pragma unroll
                            for (int i = 0; i < NU; i++) recL[col[t] * NU + i] = Lk_c[t][i];
                        }
                    }
                }
                for (int e = NU * NX + g; e < Rec<P>::RLM; e += G) recL[e] = 0.0;
                if (g == 0) {
                    double recl[Rec<P>::RLS];
#pragma unroll
                    for (int i = 0; i < Rec<P>::RLS; i++) recl[i] = (i < NU) ? lk[i] : 0.0;
                    st_rec<Rec<P>::RLS>(w.Ll[cur] + ((size_t)k * Bp + b) * Rec<P>::RLS, recl);
                }
            }
            /* ---- expected reduction (back_pass.c:204-214), every lane keeps the same running sums ---- */
#pragma unroll
            for (int i = 0; i < NU; i++) dV0 += Qu[i] * lk[i];
#pragma unroll
            for (int i = 0; i < NU; i++) {
                double acc = 0.0;
#pragma unroll
                for (int j = 0; j < NU; j++) acc += lk[j] * Quu[symtri(j, i)];
                dV1 += 0.5 * lk[i] * acc;
            }
            /* ---- value function, own columns (back_pass.c:217-241) ---- */
            double bv[NU];
#pragma unroll
            for (int r = 0; r < NU; r++) {
                double acc = 0.0;
#pragma unroll
                for (int s = 0; s < NU; s++) acc += Quu[symtri(r, s)] * lk[s];
                bv[r] = acc;
            }
            double Vx_c[CPL], Vxx_d[CPL];
#pragma unroll
            for (int t = 0; t < CPL; t++) {
                double blc[NU];
                {
                    double acc = 0.0;
#pragma unroll
                    for (int s = 0; s < NU; s++) acc += Lk_c[t][s] * bv[s];
                    double v = Qx_c[t] + acc;
#pragma unroll
                    for (int j = 0; j < NU; j++) v += Lk_c[t][j] * Qu[j];
#pragma unroll
                    for (int j = 0; j < NU; j++) v += Qxu_c[t][j] * lk[j];
                    Vx_c[t] = v;
                }
#pragma unroll
                for (int r = 0; r < NU; r++) {
                    double acc = 0.0;
#pragma unroll
                    for (int s = 0; s < NU; s++) acc += Quu[symtri(r, s)] * Lk_c[t][s];
                    blc[r] = acc;
                }
                {
                    double acc = 0.0;
#pragma unroll
                    for (int s = 0; s < NU; s++) acc += Lk_c[t][s] * blc[s];
                    double v = Qxx_d[t] + acc;
#pragma unroll
                    for (int cc = 0; cc < NU; cc++) {
                        double term = Lk_c[t][cc] * Qxu_c[t][cc];
                        term *= 2.0;
                        v += term;
                    }
                    Vxx_d[t] = v;
                }
                if (colv[t]) {
                    sts_vec<NU, ALU>(ws + WS::OFF_BL + col[t] * NU, blc);
                    ws[WS::OFF_V + col[t]] = Vx_c[t];
                    ws[WS::OFF_V + NX + (col[t] * (col[t] + 1)) / 2 + col[t]] = Vxx_d[t];
                }
            }
            __syncwarp(gmask);
#pragma unroll
            for (int t = 0; t < ROUNDS; t++) {
                const int r = orow[t], c = ocol[t];
                double Lr[NU], Lc[NU], br[NU], bc[NU], qr[NU], qc[NU];
                lds_vec<NU, ALU>(ws + WS::OFF_LK + r * NU, Lr);
                lds_vec<NU, ALU>(ws + WS::OFF_LK + c * NU, Lc);
                lds_vec<NU, ALU>(ws + WS::OFF_BL + r * NU, br);
                lds_vec<NU, ALU>(ws + WS::OFF_BL + c * NU, bc);
                lds_vec<NU, ALU>(ws + WS::OFF_QT + r * NU, qr);
                lds_vec<NU, ALU>(ws + WS::OFF_QT + c * NU, qc);
                double acc = 0.0;
#pragma unroll
                for (int s = 0; s < NU; s++) acc += Lr[s] * bc[s];
#pragma unroll
                for (int s = 0; s < NU; s++) acc += Lc[s] * br[s];
                acc *= 0.5;
                double v = Qxx_o[t] + acc;
                /* the reference's loop visits (i=r, j=c) before (i=c, j=r) for r < c */
#pragma unroll
                for (int cc = 0; cc < NU; cc++) v += Lr[cc] * qc[cc];
#pragma unroll
                for (int cc = 0; cc < NU; cc++) v += Lc[cc] * qr[cc];
                if (offv[t]) ws[WS::OFF_V + NX + (c * (c + 1)) / 2 + r] = v;
            }
            /* ---- gradient measure (back_pass.c:244-251) ---- */
            {
                double gmax = 0.0;
#pragma unroll
                for (int i = 0; i < NU; i++) {
                    const double gi = fabs(lk[i]) / (fabs(un[i]) + 1.0);
                    if (gi > gmax) gmax = gi;
                }
                g_sum += gmax;
            }
            __syncwarp(gmask);
            lds_vec<NX, true>(ws + WS::OFF_V, Vx);
            lds_vec<NQXX, (NX % 2 == 0)>(ws + WS::OFF_V + NX, Vxx);
            /* no barrier here: the next write to any of these arrays sits behind the next step's first barrier */
        }
        __syncwarp(gmask);
        if (failed) {
            if (o.bp_single) break;   /* back_pass(o) on its own: one attempt, the retry loop belongs to iLQG() */
            raise_lambda(o, lambda, dlambda);
            if (lambda > o.lambdaMax) break;
        } else {
            done = true;
        }
    }
    if (g != 0) return;
    w.n_bp[b] = n_bp;
    w.dV0[b] = dV0;
    w.dV1[b] = dV1;
    w.bp_done[b] = done ? 1 : 0;
    double g_norm = w.g_norm[b];
    if (done) {
        g_norm = g_sum / ((double)(T - 1));
        w.g_norm[b] = g_norm;
    }
    if (o.bp_single) return;
    if (g_norm < o.tolGrad && lambda < 1e-5) { /* iLQG.c:297-303 */
        lower_lambda(o, lambda, dlambda);
        finish(w, b, iter, done ? 1 : 0);
    } else if (!done) {
        finish(w, b, iter, 0);
    }
    w.lambda[b] = lambda;
    w.dlambda[b] = dlambda;
}

} /* namespace ilqg */
