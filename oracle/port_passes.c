/* oracle/port_passes.c -- TEST INFRASTRUCTURE: CPU restatement of the backward Riccati pass and the line search.
 *
 * Independent re-write of /root/reference back_pass.c:38-257 and line_search.c:33-78 (single-threaded path
 * only).  Operation order is the reference's; structure and names are ours.  Quirks kept on purpose
 * (SURVEY.md 7.4): g_norm divides by n_hor-1 (Q1); warm start of the QP from step k+1 (Q7); regType 2 as
 * written (Q11); line search accepts the FIRST alpha with z > zMin and leaves z/dcost/expected untouched
 * when every rollout fails (Q5).
 */
#include <math.h>
#include <string.h>
#include "ilqg_compat.h"

typedef struct {
    double Qx[N_X], Qu[N_U], Qxx[sizeofQxx], Quu[sizeofQuu], Qxu[sizeofQxu];
    double QuuF[sizeofQuu], Qxu_reg[sizeofQxu];
} qfun_t;

/* Q-function of step t given the value function of step t+1 (back_pass.c:80-131) */
static void assemble_q(const trajEl_t *t, const double *Vx, const double *Vxx, qfun_t *q, double *scratch)
{
    memcpy(q->Qu, t->cu, sizeof q->Qu);
    addMulVec(q->Qu, Vx, t->fu, N_X, N_U);
    memcpy(q->Qx, t->cx, sizeof q->Qx);
    addMulVec(q->Qx, Vx, t->fx, N_X, N_X);

    memcpy(q->Qxu, t->cxu, sizeof q->Qxu);
    addMul2Tri(q->Qxu, Vxx, t->fx, N_X, N_X, t->fu, N_X, N_U, scratch);
#if FULL_DDP
    for (int e = 0; e < N_X * N_U; e++) {
        double acc = 0.0;
        for (int i = 0; i < N_X; i++)
            acc += Vx[i] * t->fxu[e + i * N_X * N_U];
        q->Qxu[e] += acc;
    }
#endif
    memcpy(q->Quu, t->cuu, sizeof q->Quu);
    addSquareTri(q->Quu, Vxx, t->fu, N_X, N_U, scratch);
#if FULL_DDP
    for (int e = 0; e < sizeofQuu; e++) {
        double acc = 0.0;
        for (int i = 0; i < N_X; i++)
            acc += Vx[i] * t->fuu[e + i * sizeofQuu];
        q->Quu[e] += acc;
    }
#endif
    memcpy(q->Qxx, t->cxx, sizeof q->Qxx);
    addSquareTri(q->Qxx, Vxx, t->fx, N_X, N_X, scratch);
#if FULL_DDP
    for (int e = 0; e < sizeofQxx; e++) {
        double acc = 0.0;
        for (int i = 0; i < N_X; i++)
            acc += Vx[i] * t->fxx[e + i * sizeofQxx];
        q->Qxx[e] += acc;
    }
#endif
}

/* Levenberg-type regularisation (back_pass.c:134-159) */
static void regularise(const trajEl_t *t, qfun_t *q, int regType, double lambda)
{
    memcpy(q->QuuF, q->Quu, sizeof q->QuuF);
    memcpy(q->Qxu_reg, q->Qxu, sizeof q->Qxu_reg);
    if (regType == 2) {
        for (int j = 0; j < N_U; j++)
            for (int i = 0; i <= j; i++) {
                double acc = 0.0;
                for (int k = 0; k < N_U; k++)
                    acc += t->fu[SYMTRI_MAT_IDX(k, i)] * t->fu[SYMTRI_MAT_IDX(k, j)];
                q->QuuF[UTRI_MAT_IDX(i, j)] += acc * lambda;
            }
        for (int i = 0; i < N_X; i++)
            for (int j = 0; j < N_U; j++) {
                double acc = 0.0;
                for (int k = 0; k < N_X; k++)
                    acc += t->fx[MAT_IDX(k, i, N_X)] * t->fu[MAT_IDX(k, j, N_U)];
                q->Qxu_reg[MAT_IDX(i, j, N_X)] += acc * lambda;
            }
    }
    if (regType == 1)
        for (int i = 0; i < N_U; i++)
            q->QuuF[UTRI_MAT_IDX(i, i)] += lambda;
}

static double active_hx(const trajEl_t *t, int which, int input, int state)
{
    return (which == 1) ? t->lower_sign[input] * t->lower_hx[MAT_IDX(state, input, N_X)]
                        : t->upper_sign[input] * t->upper_hx[MAT_IDX(state, input, N_X)];
}

/* feedback gains incl. clamped rows and state-dependent constraints (back_pass.c:173-201) */
static void gains(trajEl_t *t, const qfun_t *q, const double *invHfree, const int *clamped)
{
    memset(t->L, 0, sizeof t->L);
    int fi = 0;
    for (int i = 0; i < N_U; i++) {
        if (clamped[i]) {
            for (int s = 0; s < N_X; s++)
                t->L[MAT_IDX(i, s, N_U)] -= active_hx(t, clamped[i], i, s);
            continue;
        }
        int fj = 0;
        for (int j = 0; j < N_U; j++) {
            if (!clamped[j]) {
                for (int s = 0; s < N_X; s++)
                    t->L[MAT_IDX(i, s, N_U)] -= invHfree[SYMTRI_MAT_IDX(fi, fj)] * q->Qxu_reg[MAT_IDX(s, j, N_X)];
                fj++;
            } else {
                double w = 0.0;
                int fk = 0;
                for (int k = 0; k < N_U; k++)
                    if (!clamped[k]) {
                        w -= invHfree[SYMTRI_MAT_IDX(fi, fk)] * q->QuuF[SYMTRI_MAT_IDX(k, j)];
                        fk++;
                    }
                for (int s = 0; s < N_X; s++)
                    t->L[MAT_IDX(i, s, N_U)] -= w * active_hx(t, clamped[j], j, s);
            }
        }
        fi++;
    }
}

int back_pass(tOptSet *o)
{
    const int N = o->n_hor;
    double Vx[N_X], Vxx[sizeofQxx], scratch[N_X * N_X];
    double invHfree[sizeofQuu], Ufac[sizeofQuu], Hfree[sizeofQuu];
    double grad[N_U], grad_clamped[N_U], search[N_U];
    int clamped[N_U], n_free;
    qfun_t q;
    double g_sum = 0.0;

    o->dV[0] = 0.0;
    o->dV[1] = 0.0;
    memcpy(Vx, o->nominal->f.cx, sizeof Vx);
    memcpy(Vxx, o->nominal->f.cxx, sizeof Vxx);

    for (int k = N - 1; k >= 0; k--) {
        trajEl_t *t = o->nominal->t + k;
        assemble_q(t, Vx, Vxx, &q, scratch);
        regularise(t, &q, o->regType, o->lambda);

        if (k == N - 1)
            memset(t->l, 0, sizeof t->l);
        else
            memcpy(t->l, (t + 1)->l, sizeof t->l);
        if (boxQP(q.QuuF, q.Qu, t->lower, t->upper, t->l, Hfree, Ufac, grad, grad_clamped, search, clamped,
                  &n_free, invHfree, N_U) < 1)
            return 1;

        gains(t, &q, invHfree, clamped);

        /* expected cost change terms (back_pass.c:204-214): accumulated straight into dV */
        for (int i = 0; i < N_U; i++)
            o->dV[0] += q.Qu[i] * t->l[i];
        for (int i = 0; i < N_U; i++) {
            double acc = 0.0;
            for (int j = 0; j < N_U; j++)
                acc += t->l[j] * q.Quu[SYMTRI_MAT_IDX(j, i)];
            o->dV[1] += 0.5 * t->l[i] * acc;
        }

        /* value function of step k (back_pass.c:217-241) */
        memcpy(Vx, q.Qx, sizeof Vx);
        addMul2Tri(Vx, q.Quu, t->L, N_U, N_X, t->l, N_U, 1, scratch);
        for (int i = 0; i < N_X; i++)
            for (int j = 0; j < N_U; j++)
                Vx[i] += t->L[MAT_IDX(j, i, N_U)] * q.Qu[j];
        for (int i = 0; i < N_X; i++)
            for (int j = 0; j < N_U; j++)
                Vx[i] += q.Qxu[MAT_IDX(i, j, N_X)] * t->l[j];

        memcpy(Vxx, q.Qxx, sizeof Vxx);
        addSquareTri(Vxx, q.Quu, t->L, N_U, N_X, scratch);
        for (int i = 0; i < N_X; i++)
            for (int j = 0; j < N_X; j++)
                for (int c = 0; c < N_U; c++) {
                    double term = t->L[MAT_IDX(c, i, N_U)] * q.Qxu[MAT_IDX(j, c, N_X)];
                    if (i == j)
                        term *= 2.0;
                    Vxx[SYMTRI_MAT_IDX(i, j)] += term;
                }

        /* gradient measure (back_pass.c:244-251) */
        double gmax = 0.0;
        for (int i = 0; i < N_U; i++) {
            double gi = fabs(t->l[i]) / (fabs(t->u[i]) + 1.0);
            if (gi > gmax)
                gmax = gi;
        }
        g_sum += gmax;
    }
    o->g_norm = g_sum / ((double)(o->n_hor - 1));
    return 0;
}

int line_search(tOptSet *o, int iter)
{
    double expected, z, dcost, cnew;
    int tried = 0, accepted = 0;

    for (tried = 0; tried < o->n_alpha; tried++) {
        const double alpha = o->alpha[tried];
        if (!forward_pass(o->candidates[0], o, alpha, &cnew, 0))
            continue; /* non-finite rollout: next alpha (line_search.c:55-59) */
        dcost = o->cost - cnew;
        expected = -alpha * (o->dV[0] + alpha * o->dV[1]);
        z = (expected > 0) ? dcost / expected : 0;
        if (z > o->zMin) {
            accepted = 1;
            break;
        }
    }
    if (o->log_linesearch != NULL) o->log_linesearch[iter] = tried + 1;
    if (o->log_z != NULL) o->log_z[iter] = z;
    if (o->log_cost != NULL) o->log_cost[iter] = cnew;
    o->new_cost = cnew;
    o->dcost = dcost;
    o->expected = expected;
    return accepted;
}
