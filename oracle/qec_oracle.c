/*
 * qec_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C restatement of the Metropolis-chain decoder hot path of
 * QEC-project-2020/MCMC-QEC-toric-RL (Python + numba).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may load this library, and only as the checker / CPU baseline -- never as
 * the product path (the product is mcmc-qec-toric-rl_b200/csrc, CUDA only).
 *
 * Parity pin: the reference ships no golden vectors or runnable tests
 * (SURVEY.md section 4), so this oracle is pinned against outputs of the
 * reference itself, run seeded in the build container by
 * tests/golden/make_golden.py and committed as tests/golden/ fixtures
 * (tests/test_oracle_golden.py replays every one of them through this file).
 *
 * Every function cites the reference file:line it follows
 * (paths relative to the reference checkout).
 *
 * Semantics reproduced on purpose:
 *   Q1  the reference's fast chain kernel always proposes with *planar*
 *       geometry; here the proposal geometry is an explicit argument
 *       (geom_chain) separate from the code's own geometry (geom_code).
 *   Q2  alpha/biased chains freeze the acceptance denominator per block.
 *   Q8  Ladder_alpha swaps .code and .flag but not .n_eff (mcmc_alpha.py:126-131).
 *   pow: numba lowers float**int to square-and-multiply (used by the njit
 *       fast path and _r_flip); CPython float**int calls libm pow().
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define QO_TORIC 0
#define QO_PLANAR 1
#define QO_ROTATED 2
#define QO_XZZX 3

#define QO_EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------ */
/* Random streams.  The reference uses three never-seeded MT19937      */
/* streams: numba's (NB), CPython's (PY), numpy's (NP). */
/* All three produce doubles by genrand_res53.                         */
/* ------------------------------------------------------------------ */
typedef struct qo_stream {
    uint32_t mt[624];
    int idx;
    const double *replay; /* if non-NULL: draw from this array instead */
    int64_t replay_len;
    int64_t drawn;
    double *record; /* optional log of every draw */
    int64_t record_cap;
    /* native streams (see "Native draws" below): Philox4x32-10 words instead of MT19937 doubles */
    int native;
    uint32_t ph_key[2], ph_ctr[4]; /* ctr[0] = index of the next call */
    uint32_t ph_buf[4];
    int ph_have;       /* unread words left in ph_buf */
    uint32_t bit_word; /* qo_stream_next_bit: the word being handed out bit by bit */
    int bit_have;
    uint32_t lad_step; /* native == 2: index of the next Ladder.step of this ladder (see ladder_step_native) */
} qo_stream;

static void mt_init_genrand(qo_stream *s, uint32_t seed)
{
    s->mt[0] = seed;
    for (int i = 1; i < 624; i++)
        s->mt[i] = 1812433253u * (s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) + (uint32_t)i;
    s->idx = 624;
}

static void mt_init_by_array(qo_stream *s, const uint32_t *key, int klen)
{
    mt_init_genrand(s, 19650218u);
    int i = 1, j = 0;
    int k = 624 > klen ? 624 : klen;
    for (; k; k--) {
        s->mt[i] = (s->mt[i] ^ ((s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
        i++; j++;
        if (i >= 624) { s->mt[0] = s->mt[623]; i = 1; }
        if (j >= klen) j = 0;
    }
    for (k = 623; k; k--) {
        s->mt[i] = (s->mt[i] ^ ((s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
        i++;
        if (i >= 624) { s->mt[0] = s->mt[623]; i = 1; }
    }
    s->mt[0] = 0x80000000u;
    s->idx = 624;
}

static uint32_t mt_next32(qo_stream *s)
{
    if (s->idx >= 624) {
        uint32_t *mt = s->mt;
        for (int k = 0; k < 624; k++) {
            uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
            mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        s->idx = 0;
    }
    uint32_t y = s->mt[s->idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

/* numba random.seed(s) inside njit, and np.random.seed(int): init_genrand */
QO_EXPORT qo_stream *qo_stream_mt(uint32_t seed)
{
    qo_stream *s = (qo_stream *)calloc(1, sizeof(qo_stream));
    mt_init_genrand(s, seed);
    return s;
}

/* CPython random.seed(non-negative int < 2**64): init_by_array of 32-bit limbs */
QO_EXPORT qo_stream *qo_stream_mt_pyseed(uint64_t seed)
{
    qo_stream *s = (qo_stream *)calloc(1, sizeof(qo_stream));
    uint32_t key[2] = {(uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32)};
    mt_init_by_array(s, key, key[1] ? 2 : 1);
    return s;
}

QO_EXPORT qo_stream *qo_stream_replay(const double *u, int64_t n)
{
    qo_stream *s = (qo_stream *)calloc(1, sizeof(qo_stream));
    s->replay = u;
    s->replay_len = n;
    return s;
}

/* ------------------------------------------------------------------ */
/* Native draws.  The product's kernels do not run MT19937: every chain */
/* owns a counter-based Philox4x32-10 stream (Salmon, Moraes, Dror,     */
/* Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11; the      */
/* round function and the constants below are the published ones, and   */
/* tests/test_oracle_native.py checks them against the Random123 known- */
/* answer vectors).  A native stream hands out the four 32-bit words of */
/* call c, then of call c + 1, ...  The reference's algorithm consumes  */
/* it through the same qo_stream_next(): a word w stands for the        */
/* uniform w / 2^32, so "u < threshold" decisions are the reference's   */
/* own comparisons.  Two things differ from the MT19937 schedule, both  */
/* distribution-preserving, and both are restated here so that the      */
/* oracle can be driven word for word like the device:                  */
/*   * a stabilizer proposal takes ONE word: idx = floor(u * n_stab)    */
/*     over the canonical numbering of qo_stabilizer_by_index (every    */
/*     stabilizer equiprobable, as under the reference's 3 or 5 draws); */
/*   * the rain takes one BIT per site (p = 0.5).                       */
/* ------------------------------------------------------------------ */
static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

QO_EXPORT void qo_philox4x32_10(const uint32_t *ctr, const uint32_t *key, uint32_t *out) { philox4x32_10(ctr, key, out); }

/* key = 64-bit seed; counter = (call index starting at call0, tag, id lo, id hi) */
QO_EXPORT qo_stream *qo_stream_philox(uint64_t key, uint64_t id, uint32_t tag, uint32_t call0)
{
    qo_stream *s = (qo_stream *)calloc(1, sizeof(qo_stream));
    s->native = 1;
    s->ph_key[0] = (uint32_t)key; s->ph_key[1] = (uint32_t)(key >> 32);
    s->ph_ctr[0] = call0; s->ph_ctr[1] = tag; s->ph_ctr[2] = (uint32_t)id; s->ph_ctr[3] = (uint32_t)(id >> 32);
    return s;
}

static uint32_t philox_next_word(qo_stream *s)
{
    if (s->ph_have == 0) {
        philox4x32_10(s->ph_ctr, s->ph_key, s->ph_buf);
        s->ph_ctr[0]++;
        s->ph_have = 4;
    }
    return s->ph_buf[4 - s->ph_have--];
}

/* A LADDER-native stream (native == 2) carries no position: ladder_step_native addresses every word it needs by
 * (Ladder.step index, purpose, rung, iteration), so the words do not depend on how many draws were made before.
 * key = the call's seed, id = the ladder's global index, step0 = Ladder.step calls already made on this ladder. */
QO_EXPORT qo_stream *qo_stream_ladder_native(uint64_t key, uint64_t id, uint32_t step0)
{
    qo_stream *s = qo_stream_philox(key, id, 0, 0);
    s->native = 2;
    s->lad_step = step0;
    return s;
}

/* drop the unread words of the current call: the next draw starts a new call */
QO_EXPORT void qo_stream_align(qo_stream *s) { s->ph_have = 0; s->bit_have = 0; }

/* one bit per draw, least significant bit of each word first (native rain) */
QO_EXPORT int qo_stream_next_bit(qo_stream *s)
{
    if (s->bit_have == 0) { s->bit_word = philox_next_word(s); s->bit_have = 32; s->drawn++; }
    int b = (int)(s->bit_word & 1u);
    s->bit_word >>= 1;
    s->bit_have--;
    return b;
}

QO_EXPORT void qo_stream_record(qo_stream *s, double *buf, int64_t cap)
{
    s->record = buf;
    s->record_cap = cap;
}

QO_EXPORT int64_t qo_stream_drawn(const qo_stream *s) { return s->drawn; }
QO_EXPORT void qo_stream_free(qo_stream *s) { free(s); }

QO_EXPORT double qo_stream_next(qo_stream *s)
{
    double u;
    if (s->native) {
        u = (double)philox_next_word(s) * (1.0 / 4294967296.0);
    } else if (s->replay) {
        u = (s->drawn < s->replay_len) ? s->replay[s->drawn] : 0.0;
    } else {
        uint32_t a = mt_next32(s) >> 5, b = mt_next32(s) >> 6;
        u = (a * 67108864.0 + b) / 9007199254740992.0;
    }
    if (s->record && s->drawn < s->record_cap) s->record[s->drawn] = u;
    s->drawn++;
    return u;
}

/* ------------------------------------------------------------------ */
/* Lattice helpers                                                     */
/* ------------------------------------------------------------------ */
QO_EXPORT int qo_nsites(int geom, int L) { return (geom == QO_TORIC || geom == QO_PLANAR) ? 2 * L * L : L * L; }
/* toric_model.py:8, planar_model.py:10, rotated_surface_model.py:9, xzzx_model.py:9 */
QO_EXPORT int qo_neq(int geom) { return geom == QO_TORIC ? 16 : 4; }

/* _count_errors: toric_model.py:174-176 (np.count_nonzero) */
QO_EXPORT int qo_count_errors(const uint8_t *qm, int n)
{
    int c = 0;
    for (int i = 0; i < n; i++) c += qm[i] != 0;
    return c;
}

/* chain_lengths / _count_errors_xyz: planar_model.py:101-105, 224-229 */
QO_EXPORT void qo_count_xyz(const uint8_t *qm, int n, int64_t out[3])
{
    out[0] = out[1] = out[2] = 0;
    for (int i = 0; i < n; i++)
        if (qm[i]) out[qm[i] - 1]++;
}

static inline int flip(uint8_t *q, int op)
{
    uint8_t old = *q, nw = old ^ (uint8_t)op;
    *q = nw;
    return (nw && !old) - (old && !nw);
}

#define I3(l, r, c) (((l) * L + (r)) * L + (c))
#define I2(r, c) ((r) * L + (c))

/* _apply_stabilizer, in place, returns the weight change.
 * toric_model.py:256-284, planar_model.py:291-339,
 * rotated_surface_model.py:349-392, xzzx_model.py:360-436.
 * (row, col, op) have the reference's meaning for each code; for
 * rotated/xzzx op==1 is a full and op==3 a half stabilizer. */
QO_EXPORT int qo_apply_stabilizer(int geom, int L, uint8_t *qm, int row, int col, int op)
{
    int d = 0;
    switch (geom) {
    case QO_TORIC:
        if (op == 1) {
            d += flip(&qm[I3(1, row, col)], 1);
            d += flip(&qm[I3(1, row, (col - 1 + L) % L)], 1);
            d += flip(&qm[I3(0, row, col)], 1);
            d += flip(&qm[I3(0, (row - 1 + L) % L, col)], 1);
        } else {
            d += flip(&qm[I3(1, row, col)], 3);
            d += flip(&qm[I3(0, row, col)], 3);
            d += flip(&qm[I3(0, row, (col + 1) % L)], 3);
            d += flip(&qm[I3(1, (row + 1) % L, col)], 3);
        }
        break;
    case QO_PLANAR:
        if (op == 1) {
            d += flip(&qm[I3(0, row, col)], 1);
            d += flip(&qm[I3(0, row + 1, col)], 1);
            if (col == 0) d += flip(&qm[I3(1, row, 0)], 1);
            else if (col == L - 1) d += flip(&qm[I3(1, row, col - 1)], 1);
            else { d += flip(&qm[I3(1, row, col)], 1); d += flip(&qm[I3(1, row, col - 1)], 1); }
        } else {
            d += flip(&qm[I3(0, row, col)], 3);
            d += flip(&qm[I3(0, row, col + 1)], 3);
            if (row == 0) d += flip(&qm[I3(1, 0, col)], 3);
            else if (row == L - 1) d += flip(&qm[I3(1, row - 1, col)], 3);
            else { d += flip(&qm[I3(1, row, col)], 3); d += flip(&qm[I3(1, row - 1, col)], 3); }
        }
        break;
    case QO_ROTATED:
        if (op == 1) {
            int p = ((row + col) % 2 == 0) ? 1 : 3;
            d += flip(&qm[I2(row, col)], p);
            d += flip(&qm[I2(row, col + 1)], p);
            d += flip(&qm[I2(row + 1, col)], p);
            d += flip(&qm[I2(row + 1, col + 1)], p);
        } else {
            switch (col) {
            case 0: d += flip(&qm[I2(0, 2 * row + 1)], 1); d += flip(&qm[I2(0, 2 * row + 2)], 1); break;
            case 1: d += flip(&qm[I2(2 * row + 1, L - 1)], 3); d += flip(&qm[I2(2 * row + 2, L - 1)], 3); break;
            case 2: d += flip(&qm[I2(L - 1, 2 * row)], 1); d += flip(&qm[I2(L - 1, 2 * row + 1)], 1); break;
            default: d += flip(&qm[I2(2 * row, 0)], 3); d += flip(&qm[I2(2 * row + 1, 0)], 3); break;
            }
        }
        break;
    default: /* QO_XZZX */
        if (op == 1) {
            d += flip(&qm[I2(row, col)], 1);
            d += flip(&qm[I2(row + 1, col)], 3);
            d += flip(&qm[I2(row, col + 1)], 3);
            d += flip(&qm[I2(row + 1, col + 1)], 1);
        } else {
            switch (col) {
            case 0: d += flip(&qm[I2(0, 2 * row + 1)], 3); d += flip(&qm[I2(0, 2 * row + 2)], 1); break;
            case 1: d += flip(&qm[I2(2 * row + 1, L - 1)], 1); d += flip(&qm[I2(2 * row + 2, L - 1)], 3); break;
            case 2: d += flip(&qm[I2(L - 1, 2 * row)], 1); d += flip(&qm[I2(L - 1, 2 * row + 1)], 3); break;
            default: d += flip(&qm[I2(2 * row, 0)], 3); d += flip(&qm[I2(2 * row + 1, 0)], 1); break;
            }
        }
        break;
    }
    return d;
}

/* Draw order and arithmetic of _apply_random_stabilizer.
 * toric_model.py:287-296 (3 draws), planar_model.py:342-352 (3 draws),
 * rotated_surface_model.py:395-408 and xzzx_model.py:439-452 (5 draws,
 * all five always consumed). */
/* Canonical numbering of a code's stabilizers for native proposals (a product convention, not the reference's):
 * toric  0 .. L^2-1: op 1 at (idx / L, idx % L); then op 3 likewise;
 * planar 0 .. L(L-1)-1: op 1 at (idx / L, idx % L), rows 0 .. L-2; then op 3 at (rem / (L-1), rem % (L-1));
 * rotated / XZZX 0 .. (L-1)^2-1: full plaquettes (op 1) at (idx / (L-1), idx % (L-1)); then the 2(L-1) half
 * plaquettes (op 3) as (k, side) = (rem / 4, rem % 4) -- the (rows2, cols2) of rotated_surface_model.py:400-407. */
QO_EXPORT int qo_nstab(int geom, int L)
{
    return geom == QO_TORIC ? 2 * L * L : geom == QO_PLANAR ? 2 * L * (L - 1) : L * L - 1;
}

QO_EXPORT void qo_stabilizer_by_index(int geom, int L, int idx, int *row, int *col, int *op)
{
    int nfull = geom == QO_TORIC ? L * L : geom == QO_PLANAR ? L * (L - 1) : (L - 1) * (L - 1);
    int full = idx < nfull, rem = full ? idx : idx - nfull;
    *op = full ? 1 : 3;
    if (geom == QO_TORIC) { *row = rem / L; *col = rem % L; }
    else if (geom == QO_PLANAR) { int w = full ? L : L - 1; *row = rem / w; *col = rem % w; }
    else if (full) { *row = rem / (L - 1); *col = rem % (L - 1); }
    else { *row = rem / 4; *col = rem % 4; }
}

QO_EXPORT void qo_draw_stabilizer(int geom, int L, qo_stream *nb, int *row, int *col, int *op)
{
    if (nb->native) {   /* one word: floor(w * n_stab / 2^32), exact in double */
        int idx = (int)(qo_stream_next(nb) * qo_nstab(geom, L));
        qo_stabilizer_by_index(geom, L, idx, row, col, op);
        return;
    }
    if (geom == QO_TORIC) {
        *row = (int)(qo_stream_next(nb) * L);
        *col = (int)(qo_stream_next(nb) * L);
        int o = (int)(qo_stream_next(nb) * 2);
        *op = o == 0 ? 3 : o;
    } else if (geom == QO_PLANAR) {
        int s = (int)((L - 1) * qo_stream_next(nb));
        int l = (int)(L * qo_stream_next(nb));
        if (qo_stream_next(nb) < 0.5) { *row = s; *col = l; *op = 1; }
        else { *row = l; *col = s; *op = 3; }
    } else {
        int rows = (int)((L - 1) * qo_stream_next(nb));
        int cols = (int)((L - 1) * qo_stream_next(nb));
        int rows2 = (int)(((double)(L - 1) / 2.0) * qo_stream_next(nb));
        int cols2 = (int)(4 * qo_stream_next(nb));
        double phalf = (double)(L * L - (L - 1) * (L - 1) - 1) / (double)(L * L - 1);
        if (qo_stream_next(nb) > phalf) { *row = rows; *col = cols; *op = 1; }
        else { *row = rows2; *col = cols2; *op = 3; }
    }
}

/* _apply_logical, in place.  toric_model.py:179-225 (op in {1,2}: X along
 * row X_pos; op in {2,3}: Z along column Z_pos; layer 1 transposed),
 * planar_model.py:234-268 (op in {1,3}: X on [0,X_pos,:]; {2,3}: Z on [0,:,Z_pos]),
 * rotated_surface_model.py:251-282 (op in {1,3}: X on column X_pos; {2,3}: Z on row Z_pos),
 * xzzx_model.py:279-313 (op in {1,2}: X on anti-diagonal; {2,3}: Z on diagonal). */
QO_EXPORT int qo_apply_logical(int geom, int L, uint8_t *qm, int op, int layer, int X_pos, int Z_pos)
{
    int d = 0;
    if (op == 0) return 0;
    if (geom == QO_TORIC) {
        int do_X = (op == 1 || op == 2), do_Z = (op == 3 || op == 2);
        for (int i = 0; i < L; i++) {
            if (do_X) d += flip(layer == 0 ? &qm[I3(0, X_pos, i)] : &qm[I3(1, i, X_pos)], 1);
            if (do_Z) d += flip(layer == 0 ? &qm[I3(0, i, Z_pos)] : &qm[I3(1, Z_pos, i)], 3);
        }
    } else if (geom == QO_PLANAR) {
        int do_X = (op == 1 || op == 3), do_Z = (op == 2 || op == 3);
        for (int i = 0; i < L; i++) {
            if (do_X) d += flip(&qm[I3(0, X_pos, i)], 1);
            if (do_Z) d += flip(&qm[I3(0, i, Z_pos)], 3);
        }
    } else if (geom == QO_ROTATED) {
        int do_X = (op == 1 || op == 3), do_Z = (op == 2 || op == 3);
        if (do_X) for (int i = 0; i < L; i++) d += flip(&qm[I2(i, X_pos)], 1);
        if (do_Z) for (int i = 0; i < L; i++) d += flip(&qm[I2(Z_pos, i)], 3);
    } else {
        int do_X = (op == 1 || op == 2), do_Z = (op == 3 || op == 2);
        if (do_X) for (int i = 0; i < L; i++) d += flip(&qm[I2(i, L - 1 - i)], 1);
        if (do_Z) for (int i = 0; i < L; i++) d += flip(&qm[I2(i, i)], 3);
    }
    return d;
}

/* _apply_random_logical.  toric_model.py:228-253 (both layer ops drawn first,
 * then positions per layer), planar_model.py:271-288, rotated_surface_model.py:331-346,
 * xzzx_model.py:340-357 (positions drawn with the toric op encoding). */
QO_EXPORT int qo_apply_random_logical(int geom, int L, uint8_t *qm, qo_stream *nb)
{
    if (geom == QO_TORIC) {
        int ops[2];
        ops[0] = (int)(qo_stream_next(nb) * 4);
        ops[1] = (int)(qo_stream_next(nb) * 4);
        int d = 0;
        for (int layer = 0; layer < 2; layer++) {
            int op = ops[layer], X_pos = 0, Z_pos = 0;
            if (op == 1 || op == 2) X_pos = (int)(qo_stream_next(nb) * L);
            if (op == 3 || op == 2) Z_pos = (int)(qo_stream_next(nb) * L);
            d += qo_apply_logical(geom, L, qm, op, layer, X_pos, Z_pos);
        }
        return d;
    }
    int op = (int)(qo_stream_next(nb) * 4), X_pos = 0, Z_pos = 0;
    if (op == 1 || op == 2) X_pos = (int)(qo_stream_next(nb) * L);
    if (op == 3 || op == 2) Z_pos = (int)(qo_stream_next(nb) * L);
    return qo_apply_logical(geom, L, qm, op, 0, X_pos, Z_pos);
}

/* generate_random_error from explicit uniforms (the workload step before the path, generate_data.py:57-118).
 * toric_form: Toric_code.generate_random_error(p_error), toric_model.py:15-24 -- u is np.random.uniform per layer,
 * pauli is np.random.randint(3) + 1; `error = qubits > p_error -> 0`, `no_error = qubits < p_error -> 1`, times pauli.
 * otherwise: generate_random_error(p_x, p_y, p_z) of planar_model.py:18-37 (layer 1 loses its last row and column),
 * rotated_surface_model.py:25-38, xzzx_model.py:16-29 -- one rand.random() per site, strict chained comparisons. */
QO_EXPORT void qo_generate_errors(int geom, int L, int toric_form, double p_error, double p_x, double p_y, double p_z,
                                  const double *u, const uint8_t *pauli, uint8_t *qm)
{
    int n = qo_nsites(geom, L);
    for (int i = 0; i < n; i++) {
        double r = u[i];
        int q = 0;
        if (toric_form) q = r < p_error ? pauli[i] : 0;
        else if (r < p_z) q = 3;
        else if (p_z < r && r < (p_z + p_x)) q = 1;
        else if ((p_z + p_x) < r && r < (p_z + p_x + p_y)) q = 2;
        qm[i] = (uint8_t)q;
    }
    if (geom == QO_PLANAR)
        for (int k = 0; k < L; k++) {
            qm[L * L + (L - 1) * L + k] = 0; /* qubit_matrix[1, -1, :] = 0 */
            qm[L * L + k * L + (L - 1)] = 0; /* qubit_matrix[1, :, -1] = 0 */
        }
}

/* failure rule of generate_data.py:137-201: np.argmax (np.argmin for single_temp) != eq_true; first extremum, NaN wins */
QO_EXPORT int64_t qo_count_failures(const double *distr, int n_eq, int64_t S, int use_argmin, const int32_t *eq_true, int32_t *choice)
{
    int64_t fails = 0;
    for (int64_t s = 0; s < S; s++) {
        const double *d = distr + s * n_eq;
        int best = 0;
        double bv = d[0];
        for (int e = 1; e < n_eq && bv == bv; e++)
            if (d[e] != d[e] || (use_argmin ? d[e] < bv : d[e] > bv)) { best = e; bv = d[e]; }
        if (choice) choice[s] = best;
        fails += best != eq_true[s];
    }
    return fails;
}

/* _define_equivalence_class.  toric_model.py:317-351, planar_model.py:379-390,
 * rotated_surface_model.py:411-420, xzzx_model.py:455-486. */
QO_EXPORT int qo_class(int geom, int L, const uint8_t *qm)
{
    if (geom == QO_TORIC) {
        int par[4] = {0, 0, 0, 0};
        for (int l = 0; l < 2; l++)
            for (int i = 0; i < L * L; i++) {
                uint8_t q = qm[l * L * L + i];
                par[2 * l] ^= (q == 1 || q == 2);
                par[2 * l + 1] ^= (q == 3 || q == 2);
            }
        return par[0] + 2 * par[1] + 4 * par[2] + 8 * par[3];
    }
    if (geom == QO_PLANAR) {
        int x = 0, z = 0;
        for (int i = 0; i < L; i++) {
            uint8_t a = qm[I3(0, i, 0)], b = qm[I3(0, 0, i)];
            x ^= (a == 1 || a == 2);
            z ^= (b == 3 || b == 2);
        }
        return x + 2 * z;
    }
    if (geom == QO_ROTATED) {
        int x = 0, z = 0;
        for (int i = 0; i < L; i++) {
            uint8_t a = qm[I2(0, i)], b = qm[I2(i, 0)];
            x ^= (a == 1 || a == 2);
            z ^= (b == 3 || b == 2);
        }
        return x + 2 * z;
    }
    int x = 0, z = 0;
    for (int i = 0; i < L; i++) {
        uint8_t a = qm[I2(0, i)], b = qm[I2(i, 0)];
        x ^= (a == 2) || (i % 2 == 0 ? a == 1 : a == 3);
        z ^= (b == 2) || (i % 2 == 0 ? b == 3 : b == 1);
    }
    return x ? (z ? 2 : 1) : (z ? 3 : 0);
}

/* Move a chain into class eq without changing its syndrome.
 * Toric: _to_class, toric_model.py:354-377.  Others: the
 * apply_logical(define_equivalence_class() ^ eq) route of decoders.py:556-560. */
QO_EXPORT void qo_to_class(int geom, int L, uint8_t *qm, int eq)
{
    int diff = eq ^ qo_class(geom, L, qm);
    if (geom == QO_TORIC) {
        int ops = diff ^ ((diff & 0xA) >> 1);
        qo_apply_logical(geom, L, qm, ops & 3, 0, 0, 0);
        qo_apply_logical(geom, L, qm, ops >> 2, 1, 0, 0);
    } else {
        qo_apply_logical(geom, L, qm, diff, 0, 0, 0);
    }
}

/* _apply_stabilizers_uniform ("rain"): toric_model.py:299-314,
 * planar_model.py:355-376.  np.random.rand(2,L,L) < p in C order; first-axis
 * index 0 means operator 3, index 1 operator 1; planar masks x-operators in the
 * last row and z-operators in the last column AFTER drawing all 2L^2 numbers. */
QO_EXPORT void qo_rain(int geom, int L, uint8_t *qm, qo_stream *np_, double p)
{
    for (int o = 0; o < 2; o++)
        for (int r = 0; r < L; r++)
            for (int c = 0; c < L; c++) {
                int hit = np_->native ? qo_stream_next_bit(np_) : qo_stream_next(np_) < p;   /* native: p = 0.5 only */
                if (geom == QO_PLANAR) {
                    if (o == 1 && r == L - 1) hit = 0;
                    if (o == 0 && c == L - 1) hit = 0;
                }
                if (hit) qo_apply_stabilizer(geom, L, qm, r, c, o == 0 ? 3 : 1);
            }
}

/* ------------------------------------------------------------------ */
/* Chains                                                              */
/* ------------------------------------------------------------------ */
/* numba's float64 ** int64 (square and multiply, reciprocal for negatives);
 * applies to _update_chain_fast (mcmc.py:158) and _r_flip (mcmc.py:149). */
QO_EXPORT double qo_numba_pow(double a, int64_t b)
{
    double r = 1.0;
    int inv = b < 0;
    uint64_t e = inv ? (uint64_t)(-b) : (uint64_t)b;
    if (e > 0x10000) return pow(a, (double)b);
    while (e) {
        if (e & 1) r *= a;
        e >>= 1;
        a *= a;
    }
    return inv ? 1.0 / r : r;
}

/* _update_chain_fast: mcmc.py:152-160.  Per step: proposal draws (NB), then
 * one NB accept draw, always consumed.  geom is the PROPOSAL geometry (the
 * reference hard-wires planar).  Optional per-step trace. */
QO_EXPORT void qo_update_chain_fast(int geom, int L, uint8_t *qm, double factor, int64_t iters,
                                    qo_stream *nb, int8_t *dE_out, uint8_t *acc_out)
{
    for (int64_t t = 0; t < iters; t++) {
        int row, col, op;
        qo_draw_stabilizer(geom, L, nb, &row, &col, &op);
        int d = qo_apply_stabilizer(geom, L, qm, row, col, op);
        int acc = qo_stream_next(nb) < qo_numba_pow(factor, d);
        if (!acc) qo_apply_stabilizer(geom, L, qm, row, col, op); /* undo */
        if (dE_out) dE_out[t] = (int8_t)d;
        if (acc_out) acc_out[t] = (uint8_t)acc;
    }
}

/* Chain.update_chain: mcmc.py:19-43.  p_logical != 0 (top rung): PY draw
 * chooses logical vs stabilizer; accept without a draw when p >= 0.75 or
 * dE <= 0, else PY draw against factor ** dE (CPython pow).  p_logical == 0:
 * stabilizer proposal (NB), PY accept draw always consumed. */
QO_EXPORT void qo_update_chain(int geom, int L, uint8_t *qm, double p, double p_logical, int64_t iters,
                               qo_stream *nb, qo_stream *py)
{
    int n = qo_nsites(geom, L);
    double factor = (p / 3.0) / (1.0 - p);
    uint8_t *save = (uint8_t *)malloc((size_t)n);
    for (int64_t t = 0; t < iters; t++) {
        int d;
        if (p_logical != 0) {
            memcpy(save, qm, (size_t)n);
            if (qo_stream_next(py) < p_logical) {
                d = qo_apply_random_logical(geom, L, qm, nb);
            } else {
                int row, col, op;
                qo_draw_stabilizer(geom, L, nb, &row, &col, &op);
                d = qo_apply_stabilizer(geom, L, qm, row, col, op);
            }
            if (p >= 0.75 || d <= 0) continue;
            if (!(qo_stream_next(py) < pow(factor, (double)d))) memcpy(qm, save, (size_t)n);
        } else {
            int row, col, op;
            qo_draw_stabilizer(geom, L, nb, &row, &col, &op);
            d = qo_apply_stabilizer(geom, L, qm, row, col, op);
            if (!(qo_stream_next(py) < pow(factor, (double)d))) qo_apply_stabilizer(geom, L, qm, row, col, op);
        }
    }
    free(save);
}

/* Chain_alpha / Chain_biased .update_chain: mcmc_alpha.py:27-70, mcmc_biased.py:21-59.
 * kind 0 = alpha (param a = pz_tilde, b = alpha), kind 1 = biased (a = p, b = eta).
 * pb is computed ONCE per call; every proposal is accepted with
 * PY < P(new)/pb.  n_eff (alpha only) is rewritten on accept (mcmc_alpha.py:56,70).
 * `num` is system_size**2 as in the reference (also for two-layer codes). */
QO_EXPORT void qo_update_chain_weighted(int kind, int geom, int L, uint8_t *qm, double a, double b,
                                        double p_logical, int64_t iters, qo_stream *nb, qo_stream *py,
                                        double *n_eff)
{
    int n = qo_nsites(geom, L);
    double num = (double)L * (double)L;
    double px, py_, pz;
    if (kind == 0) {
        double pz_tilde = a, alpha = b;
        double p_tilde = pz_tilde + 2 * pow(pz_tilde, alpha);
        double p = p_tilde / (1 + p_tilde);
        pz = pz_tilde * (1 - p);
        px = py_ = pow(pz_tilde, alpha) * (1 - p);
    } else {
        double p = a, eta = b;
        pz = p * eta / (eta + 1);
        px = p / (2 * (eta + 1));
        py_ = px;
    }
    int64_t c[3];
    qo_count_xyz(qm, n, c);
    double q0 = 1 - px - py_ - pz;
    double pb = pow(px, (double)c[0]) * pow(py_, (double)c[1]) * pow(pz, (double)c[2]) *
                pow(q0, num - (double)c[0] - (double)c[1] - (double)c[2]);
    uint8_t *save = (uint8_t *)malloc((size_t)n);
    for (int64_t t = 0; t < iters; t++) {
        memcpy(save, qm, (size_t)n);
        if (p_logical != 0 && qo_stream_next(py) < p_logical) {
            qo_apply_random_logical(geom, L, qm, nb);
        } else {
            int row, col, op;
            qo_draw_stabilizer(geom, L, nb, &row, &col, &op);
            qo_apply_stabilizer(geom, L, qm, row, col, op);
        }
        qo_count_xyz(qm, n, c);
        double pn = pow(px, (double)c[0]) * pow(py_, (double)c[1]) * pow(pz, (double)c[2]) *
                    pow(q0, num - (double)c[0] - (double)c[1] - (double)c[2]);
        if (qo_stream_next(py) < pn / pb) {
            if (kind == 0 && n_eff) *n_eff = (double)c[2] + b * (double)(c[0] + c[1]);
        } else {
            memcpy(qm, save, (size_t)n);
        }
    }
    free(save);
}

/* ------------------------------------------------------------------ */
/* Ladder.step on native words.  The algorithm is the reference's        */
/* (Chain / Chain_alpha / Chain_biased .update_chain on every rung, then */
/* the top-to-bottom swap sweep, mcmc.py:19-43,94-103, mcmc_alpha.py:    */
/* 27-70,117-137, mcmc_biased.py:21-59,107-124); only where each uniform */
/* comes from differs.  Word k of Philox call (c0, tag) under the         */
/* ladder's key and id, tag = purpose << 8 | rung:                        */
/*   purpose 0, c0 = s * ceil(iters / 2) + it / 2: stabilizer proposal    */
/*     (word 2 (it & 1)) and accept draw (word 2 (it & 1) + 1) of         */
/*     iteration `it` of the rung in Ladder.step number s;                */
/*   purpose 1, c0 = s * iters + it (top rung, p_logical != 0): word 0    */
/*     decides logical vs stabilizer (u < p_logical), word 1 is the       */
/*     accept draw of a logical move, word 2 carries the operators        */
/*     (bits 31:30 layer 0, bits 29:28 the toric code's layer 1);         */
/*   purpose 2, same c0: words 0..3 = X_pos, Z_pos of layer 0, then of    */
/*     layer 1 (floor(u L)); as in the reference a position is only       */
/*     drawn for operator 1 / 2 (X_pos) and 3 / 2 (Z_pos), else it is 0;  */
/*   purpose 3, c0 = s: word 0 is the swap draw of pair (rung, rung + 1). */
/* Weighted chains accept iff u * pb < pn (the product's form of          */
/* u < pn / pb: one rounding instead of a division), pb frozen per block. */
/* ------------------------------------------------------------------ */
static void lad_words(const qo_stream *st, uint32_t c0, int purpose, int rung, uint32_t w[4])
{
    uint32_t ctr[4] = {c0, (uint32_t)(purpose << 8 | rung), st->ph_ctr[2], st->ph_ctr[3]};
    philox4x32_10(ctr, st->ph_key, w);
}

static void ladder_step_native(int kind, int geom, int L, int Nc, uint8_t *qm, const double *ladder,
                               const double *diff, double param_b, double p_logical, int32_t *flags,
                               double *n_eff, int64_t *tops0, int64_t iters, qo_stream *st)
{
    const int n = qo_nsites(geom, L), nstab = qo_nstab(geom, L);
    const uint32_t s = st->lad_step++, H = (uint32_t)((iters + 1) / 2);
    const double U = 1.0 / 4294967296.0, num = (double)L * (double)L;
    uint8_t *save = (uint8_t *)malloc((size_t)n);
    for (int i = 0; i < Nc; i++) {
        uint8_t *q = qm + (size_t)i * n;
        const int is_top = (i == Nc - 1) && p_logical != 0;
        double factor = 0, px = 0, py_ = 0, pz = 0, q0 = 0, pb = 0, p = ladder[i];
        int64_t c[3];
        if (kind == 0) {
            factor = (p / 3.0) / (1.0 - p);
        } else {
            if (kind == 1) {
                double pz_tilde = ladder[i], p_tilde = pz_tilde + 2 * pow(pz_tilde, param_b), pp = p_tilde / (1 + p_tilde);
                pz = pz_tilde * (1 - pp);
                px = py_ = pow(pz_tilde, param_b) * (1 - pp);
            } else {
                pz = p * param_b / (param_b + 1);
                px = py_ = p / (2 * (param_b + 1));
            }
            q0 = 1 - px - py_ - pz;
            qo_count_xyz(q, n, c);
            pb = pow(px, (double)c[0]) * pow(py_, (double)c[1]) * pow(pz, (double)c[2]) * pow(q0, num - (double)c[0] - (double)c[1] - (double)c[2]);
        }
        for (int64_t it = 0; it < iters; it++) {
            uint32_t w[4], t[4], pos[4];
            lad_words(st, s * H + (uint32_t)(it >> 1), 0, i, w);
            double u_acc = (double)w[2 * (it & 1) + 1] * U;
            int logical = 0, d;
            memcpy(save, q, (size_t)n);
            if (is_top) {
                lad_words(st, s * (uint32_t)iters + (uint32_t)it, 1, i, t);
                logical = (double)t[0] * U < p_logical;
            }
            if (logical) {
                u_acc = (double)t[1] * U;
                lad_words(st, s * (uint32_t)iters + (uint32_t)it, 2, i, pos);
                d = 0;
                for (int layer = 0; layer < (geom == QO_TORIC ? 2 : 1); layer++) {
                    int op = (int)((t[2] >> (30 - 2 * layer)) & 3u);
                    int X_pos = (op == 1 || op == 2) ? (int)(((uint64_t)pos[2 * layer] * (uint64_t)L) >> 32) : 0;
                    int Z_pos = (op == 3 || op == 2) ? (int)(((uint64_t)pos[2 * layer + 1] * (uint64_t)L) >> 32) : 0;
                    d += qo_apply_logical(geom, L, q, op, layer, X_pos, Z_pos);
                }
            } else {
                int row, col, op;
                qo_stabilizer_by_index(geom, L, (int)(((uint64_t)w[2 * (it & 1)] * (uint64_t)nstab) >> 32), &row, &col, &op);
                d = qo_apply_stabilizer(geom, L, q, row, col, op);
            }
            int acc;
            if (kind == 0) {
                if (is_top && (p >= 0.75 || d <= 0)) acc = 1;   /* mcmc.py:30-31: no draw needed */
                else acc = u_acc < pow(factor, (double)d);
            } else {
                qo_count_xyz(q, n, c);
                double pn = pow(px, (double)c[0]) * pow(py_, (double)c[1]) * pow(pz, (double)c[2]) *
                            pow(q0, num - (double)c[0] - (double)c[1] - (double)c[2]);
                acc = u_acc * pb < pn;
                if (acc && kind == 1) n_eff[i] = (double)c[2] + param_b * (double)(c[0] + c[1]);
            }
            if (!acc) memcpy(q, save, (size_t)n);
        }
    }
    for (int i = Nc - 2; i >= 0; i--) {
        uint8_t *lo = qm + (size_t)i * n, *hi = lo + n;
        uint32_t w[4];
        lad_words(st, s, 3, i, w);
        const double u = (double)w[0] * U;
        int swap;
        if (kind == 1) {
            swap = u < pow(ladder[i] / ladder[i + 1], n_eff[i + 1] - n_eff[i]);
        } else {
            int ne_lo = qo_count_errors(lo, n), ne_hi = qo_count_errors(hi, n);
            if (kind == 0 && ne_hi < ne_lo) swap = 1;
            else swap = u < qo_numba_pow(diff[i], (int64_t)ne_hi - ne_lo);
        }
        if (swap) {
            memcpy(save, lo, (size_t)n); memcpy(lo, hi, (size_t)n); memcpy(hi, save, (size_t)n);
            int32_t f = flags[i]; flags[i] = flags[i + 1]; flags[i + 1] = f;
        }
    }
    free(save);
    flags[Nc - 1] = 1;
    if (flags[0] == 1) { (*tops0)++; flags[0] = 0; }
}

/* ------------------------------------------------------------------ */
/* Ladders (parallel tempering)                                        */
/* ------------------------------------------------------------------ */
/* kind: 0 depolarizing (mcmc.py:49-103), 1 alpha (mcmc_alpha.py:77-137),
 * 2 biased (mcmc_biased.py:66-124).  qm: [Nc][n_sites] rung states, rung 0
 * = coldest.  ladder[]: p per rung (or pz_tilde per rung for alpha);
 * diff[]: p_diff (unused for alpha).  flags[Nc], n_eff[Nc] (alpha), tops0.
 * One call = Ladder.step(iters): every rung bottom-to-top runs `iters`
 * steps; then the sequential top-to-bottom swap sweep. */
QO_EXPORT void qo_ladder_step(int kind, int geom, int L, int Nc, uint8_t *qm, const double *ladder,
                              const double *diff, double param_b, double p_logical, int32_t *flags,
                              double *n_eff, int64_t *tops0, int64_t iters, qo_stream *nb, qo_stream *py)
{
    int n = qo_nsites(geom, L);
    if (nb->native == 2) {   /* native words: same algorithm, positional draws */
        ladder_step_native(kind, geom, L, Nc, qm, ladder, diff, param_b, p_logical, flags, n_eff, tops0, iters, nb);
        return;
    }
    for (int i = 0; i < Nc; i++) {
        double pl = (i == Nc - 1) ? p_logical : 0.0;
        uint8_t *q = qm + (size_t)i * n;
        if (kind == 0) qo_update_chain(geom, L, q, ladder[i], pl, iters, nb, py);
        else if (kind == 1) qo_update_chain_weighted(0, geom, L, q, ladder[i], param_b, pl, iters, nb, py, &n_eff[i]);
        else qo_update_chain_weighted(1, geom, L, q, ladder[i], param_b, pl, iters, nb, py, NULL);
    }
    uint8_t *tmp = (uint8_t *)malloc((size_t)n);
    for (int i = Nc - 2; i >= 0; i--) {
        uint8_t *lo = qm + (size_t)i * n, *hi = lo + n;
        int swap;
        if (kind == 1) {
            /* mcmc_alpha.py:117-123: PY draw always; float exponent; n_eff stays with the rung */
            swap = qo_stream_next(py) < pow(ladder[i] / ladder[i + 1], n_eff[i + 1] - n_eff[i]);
        } else {
            int ne_lo = qo_count_errors(lo, n), ne_hi = qo_count_errors(hi, n);
            if (kind == 0 && ne_hi < ne_lo) swap = 1; /* mcmc.py:144-149: no draw */
            else swap = qo_stream_next(nb) < qo_numba_pow(diff[i], (int64_t)ne_hi - ne_lo); /* mcmc_biased.py:154-156 always draws */
        }
        if (swap) {
            memcpy(tmp, lo, (size_t)n); memcpy(lo, hi, (size_t)n); memcpy(hi, tmp, (size_t)n);
            int32_t f = flags[i]; flags[i] = flags[i + 1]; flags[i + 1] = f;
        }
    }
    free(tmp);
    flags[Nc - 1] = 1;
    if (flags[0] == 1) { (*tops0)++; flags[0] = 0; }
}

/* ------------------------------------------------------------------ */
/* Exact distinct-chain set (stands in for the dict keyed by            */
/* hash(qubit_matrix.tobytes()), decoders.py:251-254; exact compare     */
/* of the bytes, so no fingerprint collisions).                         */
/* ------------------------------------------------------------------ */
typedef struct qo_set {
    int n;             /* bytes per state */
    int64_t cap, cnt;  /* slots, entries */
    int64_t *slot;     /* entry index + 1, 0 = empty */
    uint8_t *arena;    /* cnt * n bytes */
    double *val;       /* optional payload per entry (3 doubles) */
    int64_t acap;
} qo_set;

static uint64_t fnv1a(const uint8_t *p, int n)
{
    uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ull; }
    h ^= h >> 29; h *= 0xbf58476d1ce4e5b9ull; h ^= h >> 32;
    return h;
}

QO_EXPORT qo_set *qo_set_new(int n)
{
    qo_set *s = (qo_set *)calloc(1, sizeof(qo_set));
    s->n = n;
    s->cap = 1024;
    s->slot = (int64_t *)calloc((size_t)s->cap, sizeof(int64_t));
    s->acap = 512;
    s->arena = (uint8_t *)malloc((size_t)s->acap * n);
    s->val = (double *)malloc((size_t)s->acap * 3 * sizeof(double));
    return s;
}

QO_EXPORT void qo_set_free(qo_set *s)
{
    if (!s) return;
    free(s->slot); free(s->arena); free(s->val); free(s);
}

QO_EXPORT int64_t qo_set_size(const qo_set *s) { return s->cnt; }
QO_EXPORT const double *qo_set_vals(const qo_set *s) { return s->val; }

static void set_grow(qo_set *s)
{
    int64_t ncap = s->cap * 2;
    int64_t *ns = (int64_t *)calloc((size_t)ncap, sizeof(int64_t));
    for (int64_t e = 0; e < s->cnt; e++) {
        uint64_t h = fnv1a(s->arena + e * s->n, s->n) & (uint64_t)(ncap - 1);
        while (ns[h]) h = (h + 1) & (uint64_t)(ncap - 1);
        ns[h] = e + 1;
    }
    free(s->slot);
    s->slot = ns;
    s->cap = ncap;
}

/* returns entry index; *is_new set to 1 when inserted */
QO_EXPORT int64_t qo_set_add(qo_set *s, const uint8_t *state, int *is_new)
{
    if ((s->cnt + 1) * 2 > s->cap) set_grow(s);
    uint64_t h = fnv1a(state, s->n) & (uint64_t)(s->cap - 1);
    while (s->slot[h]) {
        int64_t e = s->slot[h] - 1;
        if (memcmp(s->arena + e * s->n, state, (size_t)s->n) == 0) { *is_new = 0; return e; }
        h = (h + 1) & (uint64_t)(s->cap - 1);
    }
    if (s->cnt == s->acap) {
        s->acap *= 2;
        s->arena = (uint8_t *)realloc(s->arena, (size_t)s->acap * s->n);
        s->val = (double *)realloc(s->val, (size_t)s->acap * 3 * sizeof(double));
    }
    memcpy(s->arena + s->cnt * s->n, state, (size_t)s->n);
    s->slot[h] = s->cnt + 1;
    *is_new = 1;
    return s->cnt++;
}

/* union: add every entry of src into dst (dict.update, decoders.py:313-314) */
QO_EXPORT void qo_set_update(qo_set *dst, const qo_set *src)
{
    for (int64_t e = 0; e < src->cnt; e++) {
        int nw;
        int64_t k = qo_set_add(dst, src->arena + e * src->n, &nw);
        memcpy(dst->val + 3 * k, src->val + 3 * e, 3 * sizeof(double));
    }
}

/* ------------------------------------------------------------------ */
/* Droplets                                                            */
/* ------------------------------------------------------------------ */
/* STDC_droplet: decoders.py:236-265.  Returns the droplet's set of distinct
 * chains (payload val[0] = length).  geom_code drives the rain
 * (code.apply_stabilizers_uniform), geom_chain the proposals. */
QO_EXPORT qo_set *qo_stdc_droplet(int geom_code, int geom_chain, int L, uint8_t *qm, double factor,
                                  int64_t steps, int64_t iters, int randomize, double conv_mult,
                                  qo_stream *nb, qo_stream *np_, int64_t *steps_done)
{
    int n = qo_nsites(geom_code, L);
    qo_set *samples = qo_set_new(n);
    double stop = (double)steps;
    int shortest = 2 * L * L;
    if (randomize) qo_rain(geom_code, L, qm, np_, 0.5);
    int64_t step = 0;
    for (; step < steps; step++) {
        qo_update_chain_fast(geom_chain, L, qm, factor, iters, nb, NULL, NULL);
        int nw;
        int64_t e = qo_set_add(samples, qm, &nw);
        if (nw) {
            int length = qo_count_errors(qm, n);
            samples->val[3 * e] = length;
            if (conv_mult != 0 && length <= shortest) { shortest = length; stop = step * conv_mult; }
        }
        if (conv_mult != 0 && step >= stop && step * 100 >= steps) { step++; break; }
    }
    if (steps_done) *steps_done = step;
    return samples;
}

/* STDC: decoders.py:268-322.  qm_init is [n_eq][n_sites] (already moved to
 * each class by the caller, or the per-class list the reference accepts).
 * nb/np_ are per (class, droplet) stream pointers, [n_eq*droplets]; callers
 * alias them to mimic the reference's single global streams (droplets == 1).
 * Outputs: eqdistr[n_eq] (percent), optional distinct[n_eq] and
 * N_hist[n_eq][n_sites+1] (distinct chains per length). */
QO_EXPORT void qo_stdc(int geom_code, int geom_chain, int L, int n_eq, const uint8_t *qm_init,
                       double p_error, double p_sampling, int droplets, int64_t steps, int64_t iters,
                       int randomize, double conv_mult, qo_stream **nb, qo_stream **np_,
                       double *eqdistr, int64_t *distinct, int64_t *N_hist)
{
    int n = qo_nsites(geom_code, L);
    double factor = (p_sampling / 3.0) / (1.0 - p_sampling);
    double beta = -log((p_error / 3) / (1 - p_error));
    uint8_t *qm = (uint8_t *)malloc((size_t)n);
    double total = 0;
    if (N_hist) memset(N_hist, 0, sizeof(int64_t) * (size_t)n_eq * (n + 1));
    for (int eq = 0; eq < n_eq; eq++) {
        qo_set *all = qo_set_new(n);
        for (int d = 0; d < droplets; d++) {
            memcpy(qm, qm_init + (size_t)eq * n, (size_t)n);
            qo_set *s = qo_stdc_droplet(geom_code, geom_chain, L, qm, factor, steps, iters, randomize,
                                        conv_mult, nb[eq * droplets + d], np_[eq * droplets + d], NULL);
            qo_set_update(all, s);
            qo_set_free(s);
        }
        double z = 0;
        for (int64_t e = 0; e < all->cnt; e++) {
            z += exp(-beta * all->val[3 * e]);
            if (N_hist) N_hist[(size_t)eq * (n + 1) + (int)all->val[3 * e]]++;
        }
        eqdistr[eq] = z;
        total += z;
        if (distinct) distinct[eq] = all->cnt;
        qo_set_free(all);
    }
    for (int eq = 0; eq < n_eq; eq++) eqdistr[eq] = eqdistr[eq] / total * 100;
    free(qm);
}

/* single_temp: decoders.py:108-135.  mean over the first max_iters-1 samples
 * of the per-class chain length (np.average(nbr_errors_chain[eq, :j]) with
 * j = max_iters-1). */
QO_EXPORT void qo_single_temp(int geom_code, int geom_chain, int L, int n_eq, const uint8_t *qm_init,
                              double p, int64_t max_iters, int64_t iters, qo_stream **nb, double *mean_out)
{
    int n = qo_nsites(geom_code, L);
    double factor = (p / 3.0) / (1.0 - p);
    uint8_t *qm = (uint8_t *)malloc((size_t)n);
    for (int eq = 0; eq < n_eq; eq++) {
        memcpy(qm, qm_init + (size_t)eq * n, (size_t)n);
        double sum = 0;
        for (int64_t j = 0; j < max_iters; j++) {
            qo_update_chain_fast(geom_chain, L, qm, factor, iters, nb[eq], NULL, NULL);
            if (j < max_iters - 1) sum += qo_count_errors(qm, n);
        }
        mean_out[eq] = sum / (double)(max_iters - 1);
    }
    free(qm);
}

/* STRC_droplet: decoders.py:745-832.  Outputs per droplet:
 *   m_hist[n_sites+1]   len_counts (visits per length, repeats included)
 *   shortest, next_shortest (max_length when unseen)
 *   returns the set of distinct chains (val[0] = length); the distinct
 *   shortest / next-shortest sets are the entries of that length. */
QO_EXPORT qo_set *qo_strc_droplet(int geom_code, int geom_chain, int L, uint8_t *qm, double factor,
                                  int64_t steps, int64_t iters, int randomize, double conv_mult,
                                  qo_stream *nb, qo_stream *np_, int64_t *m_hist, int *shortest_out,
                                  int *next_out)
{
    int n = qo_nsites(geom_code, L);
    int max_length = 2 * L * L;
    qo_set *uniq = qo_set_new(n);
    memset(m_hist, 0, sizeof(int64_t) * (size_t)(n + 1));
    int shortest = max_length, next_shortest = max_length;
    double stop = (double)steps;
    if (randomize) qo_rain(geom_code, L, qm, np_, 0.5);
    for (int64_t step = 0; step < steps; step++) {
        qo_update_chain_fast(geom_chain, L, qm, factor, iters, nb, NULL, NULL);
        int nw;
        int64_t e = qo_set_add(uniq, qm, &nw);
        if (!nw) {
            m_hist[(int)uniq->val[3 * e]]++;
        } else {
            int length = qo_count_errors(qm, n);
            uniq->val[3 * e] = length;
            if (m_hist[length] > 0) {
                m_hist[length]++;
                if (length == shortest && conv_mult != 0) stop = step * conv_mult;
            } else {
                m_hist[length] = 1;
                if (length < shortest) {
                    next_shortest = shortest;
                    shortest = length;
                    if (conv_mult != 0) stop = step * conv_mult;
                } else if (length < next_shortest) {
                    next_shortest = length;
                }
            }
        }
        if (conv_mult != 0 && step >= stop && step * 100 >= steps) break;
    }
    *shortest_out = shortest;
    *next_out = next_shortest;
    return uniq;
}

static int64_t count_len(const qo_set *s, int length)
{
    int64_t c = 0;
    for (int64_t e = 0; e < s->cnt; e++) c += ((int)s->val[3 * e] == length);
    return c;
}

/* STRC: decoders.py:835-949, including the order-dependent droplet merge
 * (:882-928).  Optional outputs m_hist[n_eq][n_sites+1], short_info[n_eq][4] =
 * (shortest, next_shortest, #distinct shortest, #distinct next-shortest). */
QO_EXPORT void qo_strc(int geom_code, int geom_chain, int L, int n_eq, const uint8_t *qm_init,
                       double p_error, double p_sampling, int droplets, int64_t steps, int64_t iters,
                       int randomize, double conv_mult, qo_stream **nb, qo_stream **np_,
                       double *eqdistr, int64_t *m_hist_out, int64_t *short_info)
{
    int n = qo_nsites(geom_code, L);
    int max_length = 2 * L * L;
    double factor = (p_sampling / 3.0) / (1.0 - p_sampling);
    double beta_error = -log((p_error / 3) / (1 - p_error));
    double beta_sampling = -log((p_sampling / 3) / (1 - p_sampling));
    double d_beta = beta_sampling - beta_error;
    uint8_t *qm = (uint8_t *)malloc((size_t)n);
    int64_t *mh = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + 1) * droplets);
    int *sh = (int *)malloc(sizeof(int) * (size_t)droplets * 2);
    qo_set **sets = (qo_set **)malloc(sizeof(qo_set *) * (size_t)droplets);
    double total = 0;
    for (int eq = 0; eq < n_eq; eq++) {
        for (int d = 0; d < droplets; d++) {
            memcpy(qm, qm_init + (size_t)eq * n, (size_t)n);
            sets[d] = qo_strc_droplet(geom_code, geom_chain, L, qm, factor, steps, iters, randomize, conv_mult,
                                      nb[eq * droplets + d], np_[eq * droplets + d], mh + (size_t)d * (n + 1),
                                      &sh[2 * d], &sh[2 * d + 1]);
        }
        int shortest, next_shortest;
        int64_t *len_counts = (int64_t *)calloc((size_t)(n + 1), sizeof(int64_t));
        qo_set *s0 = qo_set_new(n), *s1 = qo_set_new(n);
        if (droplets == 1) {
            shortest = sh[0];
            next_shortest = sh[1];
        } else {
            shortest = max_length;
            next_shortest = max_length;
            for (int d = 0; d < droplets; d++) {
                if (sh[2 * d] < shortest) { next_shortest = shortest; shortest = sh[2 * d]; }
                if (sh[2 * d + 1] < next_shortest) next_shortest = sh[2 * d + 1];
            }
        }
        for (int d = 0; d < droplets; d++) {
            for (int l = 0; l <= n; l++) len_counts[l] += mh[(size_t)d * (n + 1) + l];
            int use0_as0 = droplets == 1 || sh[2 * d] == shortest;
            int use0_as1 = droplets > 1 && sh[2 * d] == next_shortest;
            int use1_as1 = droplets == 1 || sh[2 * d + 1] == next_shortest;
            for (int64_t e = 0; e < sets[d]->cnt; e++) {
                int len = (int)sets[d]->val[3 * e], nw;
                const uint8_t *st = sets[d]->arena + e * n;
                if (len == sh[2 * d] && use0_as0) qo_set_add(s0, st, &nw);
                if (len == sh[2 * d] && use0_as1) qo_set_add(s1, st, &nw);
                if (len == sh[2 * d + 1] && sh[2 * d + 1] != max_length && use1_as1) qo_set_add(s1, st, &nw);
            }
        }
        /* the reference's sets keep a 'temp' placeholder entry until first cleared
         * (decoders.py:750): set 1 has exactly one entry when no second length was seen */
        double shortest_count = (double)s0->cnt;
        double next_count = (double)s1->cnt;
        double shortest_fraction = shortest_count / (double)len_counts[shortest];
        double mean_fraction;
        if (next_shortest != max_length) {
            double next_fraction = next_count / (double)len_counts[next_shortest];
            mean_fraction = 0.5 * (shortest_fraction + next_fraction * exp(-beta_sampling * (next_shortest - shortest)));
        } else {
            mean_fraction = shortest_fraction;
        }
        double z = 0;
        for (int l = 0; l <= n; l++)
            if (len_counts[l]) z += (double)len_counts[l] * exp(-beta_sampling * shortest + d_beta * l);
        z *= mean_fraction;
        eqdistr[eq] = z;
        total += z;
        if (m_hist_out) memcpy(m_hist_out + (size_t)eq * (n + 1), len_counts, sizeof(int64_t) * (size_t)(n + 1));
        if (short_info) {
            short_info[4 * eq] = shortest; short_info[4 * eq + 1] = next_shortest;
            short_info[4 * eq + 2] = s0->cnt; short_info[4 * eq + 3] = s1->cnt;
        }
        for (int d = 0; d < droplets; d++) qo_set_free(sets[d]);
        qo_set_free(s0); qo_set_free(s1);
        free(len_counts);
        (void)count_len;
    }
    for (int eq = 0; eq < n_eq; eq++) eqdistr[eq] = eqdistr[eq] / total * 100;
    free(qm); free(mh); free(sh); free(sets);
}

/* STDC_droplet_alpha + STDC_Nall_n_alpha (EWD-style): decoders.py:510-581.
 * Chain_alpha.update_chain(5) per sample (slow path, Q2), distinct set of
 * eff_len = nz + alpha*(nx+ny); Z = sum exp(log(pz_tilde) * eff_len). */
QO_EXPORT void qo_stdc_alpha(int geom, int L, int n_eq, const uint8_t *qm_init, double pz_tilde_sampling,
                             double alpha, double pz_tilde, int64_t steps, int64_t iters, qo_stream *nb,
                             qo_stream *py, double *eqdistr, int64_t *distinct)
{
    int n = qo_nsites(geom, L);
    double beta = -log(pz_tilde);
    uint8_t *qm = (uint8_t *)malloc((size_t)n);
    double total = 0;
    for (int eq = 0; eq < n_eq; eq++) {
        memcpy(qm, qm_init + (size_t)eq * n, (size_t)n);
        qo_set *seen = qo_set_new(n);
        double z = 0, n_eff = 0;
        for (int64_t s = 0; s < steps; s++) {
            if (nb->native == 2) {   /* native words: the chain is a one-rung alpha ladder (no swap partner) */
                int32_t fl = 1;
                int64_t t0 = 0;
                qo_ladder_step(1, geom, L, 1, qm, &pz_tilde_sampling, NULL, alpha, 0.0, &fl, &n_eff, &t0, iters, nb, py);
            } else {
                qo_update_chain_weighted(0, geom, L, qm, pz_tilde_sampling, alpha, 0.0, iters, nb, py, &n_eff);
            }
            int nw;
            qo_set_add(seen, qm, &nw);
            if (nw) {
                int64_t c[3];
                qo_count_xyz(qm, n, c);
                z += exp(-beta * ((double)c[2] + alpha * (double)(c[0] + c[1])));
            }
        }
        eqdistr[eq] = z;
        total += z;
        if (distinct) distinct[eq] = seen->cnt;
        qo_set_free(seen);
    }
    for (int eq = 0; eq < n_eq; eq++) eqdistr[eq] = eqdistr[eq] / total * 100;
    free(qm);
}

/* PTEQ / PTEQ_biased / PTEQ_alpha: decoders.py:25-105,
 * decoders_biasednoise.py:28-90,175-237.  kind as in qo_ladder_step.
 * qm0: single init state [n_sites] (deep-copied to every rung).
 * Returns steps executed; eq_counts[n_eq] = cumulative class counts at
 * since_burn; *since_burn_out; result percent = eq_counts/(since_burn+1)*100 -> uint8. */
/* qo_pteq_ex additionally keeps PTEQ_alpha_with_shortest's bookkeeping (decoders_biasednoise.py:114-146) when
 * short_len != NULL: per class the smallest recorded bottom-rung value, how many samples hit it (short_n) and how
 * many distinct bottom-rung states were seen at it (short_unique). */
static int qo_fast_windows = 0;
QO_EXPORT void qo_set_fast_windows(int on) { qo_fast_windows = on; }

QO_EXPORT int64_t qo_pteq_ex(int kind, int geom, int L, int Nc, const uint8_t *qm0, const double *ladder,
                             const double *diff, double param_b, double p_logical, int SEQ, int TOPS,
                             int tops_burn, double eps, int64_t steps, int64_t iters, int use_conv,
                             qo_stream *nb, qo_stream *py, int64_t *eq_counts, int64_t *since_burn_out,
                             int64_t *tops0_out, uint8_t *percent_out, double *short_len, int64_t *short_n,
                             int64_t *short_unique)
{
    qo_set *uniq[16] = {0};
    if (short_len)
        for (int e = 0; e < qo_neq(geom); e++) { short_len[e] = 100000; short_n[e] = 0; uniq[e] = qo_set_new(qo_nsites(geom, L)); }
    int n = qo_nsites(geom, L), n_eq = qo_neq(geom);
    uint8_t *qm = (uint8_t *)malloc((size_t)n * Nc);
    int32_t *flags = (int32_t *)calloc((size_t)Nc, sizeof(int32_t));
    double *n_eff = (double *)calloc((size_t)Nc, sizeof(double));
    int64_t c[3];
    qo_count_xyz(qm0, n, c);
    for (int i = 0; i < Nc; i++) {
        memcpy(qm + (size_t)i * n, qm0, (size_t)n);
        n_eff[i] = (double)c[2] + param_b * (double)(c[0] + c[1]); /* mcmc_alpha.py:18-22 */
    }
    flags[Nc - 1] = 1;
    int64_t tops0 = 0, since_burn = 0, resulting_burn_in = 0, conv_start = 0, conv_streak = 0;
    int64_t w_a2 = 0, w_b2 = 0, w_a4 = 0, w_b4 = 0;   /* window edges [a2, b2) and [a4, b4) of the running sums */
    double w_S2 = 0, w_S4 = 0;
    int64_t hcap = 1024;
    double *hist = (double *)calloc((size_t)hcap, sizeof(double));
    memset(eq_counts, 0, sizeof(int64_t) * (size_t)n_eq);
    int64_t step = 0;
    for (; step < steps; step++) {
        qo_ladder_step(kind, geom, L, Nc, qm, ladder, diff, param_b, p_logical, flags, n_eff, &tops0, iters, nb, py);
        int cur = qo_class(geom, L, qm);
        if (tops0 >= tops_burn) {
            since_burn = step - resulting_burn_in;
            eq_counts[cur]++;
            if (since_burn >= hcap) {
                hist = (double *)realloc(hist, sizeof(double) * (size_t)hcap * 2);
                memset(hist + hcap, 0, sizeof(double) * (size_t)hcap);
                hcap *= 2;
            }
            hist[since_burn] = (kind == 1) ? n_eff[0] : (double)qo_count_errors(qm, n);
            if (short_len) {
                double v = hist[since_burn];
                int nw;
                if (v < short_len[cur]) {
                    short_n[cur] = 1;
                    short_len[cur] = v;
                    qo_set_free(uniq[cur]);
                    uniq[cur] = qo_set_new(n);
                    qo_set_add(uniq[cur], qm, &nw);
                } else if (v == short_len[cur]) {
                    short_n[cur]++;
                    qo_set_add(uniq[cur], qm, &nw);
                }
            }
        } else {
            resulting_burn_in++;
        }
        /* The two windows of conv_crit_error_based_PT as running sums.  For the depolarizing and biased ladders the history
           holds integers (error counts), whose sums are exact in double in any order: the running sums ARE the sums the
           reference recomputes from scratch at every step (decoders.py:93-105), at O(1) instead of O(history) per step.
           The alpha ladders' history holds n_eff = nz + alpha * nxy (not integers): they keep the literal recomputation
           unless qo_set_fast_windows(1) was called (long measurement runs only; the sums then differ in the last bits). */
        if (use_conv && tops0 >= tops_burn && (kind != 1 || qo_fast_windows)) {
            int64_t l = since_burn + 1;
            while (w_b2 < l / 2) w_S2 += hist[w_b2++];
            while (w_a2 < l / 4) w_S2 -= hist[w_a2++];
            while (w_b4 < l) w_S4 += hist[w_b4++];
            while (w_a4 < 3 * l / 4) w_S4 -= hist[w_a4++];
        }
        if (use_conv && tops0 >= TOPS) {
            /* conv_crit_error_based_PT: decoders.py:93-105 */
            int64_t l = since_burn + 1;
            double q2 = 0, q4 = 0;
            if ((kind != 1 || qo_fast_windows) && tops0 >= tops_burn) {
                q2 = w_S2;
                q4 = w_S4;
            } else {
                for (int64_t k = l / 4; k < l / 2; k++) q2 += hist[k];
                for (int64_t k = 3 * l / 4; k < l; k++) q4 += hist[k];
            }
            q2 /= (double)(l / 2 - l / 4);
            q4 /= (double)(l - 3 * l / 4);
            double err = fabs(q2 - q4);
            if (err < eps) {
                if (conv_streak >= SEQ) { step++; break; }
                conv_streak = tops0 - conv_start;
            } else {
                conv_streak = 0;
                conv_start = tops0;
            }
        }
    }
    if (since_burn_out) *since_burn_out = since_burn;
    if (tops0_out) *tops0_out = tops0;
    if (percent_out)
        for (int e = 0; e < n_eq; e++)
            percent_out[e] = (uint8_t)((double)eq_counts[e] / (double)(since_burn + 1) * 100);
    if (short_len)
        for (int e = 0; e < n_eq; e++) { short_unique[e] = uniq[e]->cnt; qo_set_free(uniq[e]); }
    free(qm); free(flags); free(n_eff); free(hist);
    return step;
}

QO_EXPORT int64_t qo_pteq(int kind, int geom, int L, int Nc, const uint8_t *qm0, const double *ladder,
                          const double *diff, double param_b, double p_logical, int SEQ, int TOPS,
                          int tops_burn, double eps, int64_t steps, int64_t iters, int use_conv,
                          qo_stream *nb, qo_stream *py, int64_t *eq_counts, int64_t *since_burn_out,
                          int64_t *tops0_out, uint8_t *percent_out)
{
    return qo_pteq_ex(kind, geom, L, Nc, qm0, ladder, diff, param_b, p_logical, SEQ, TOPS, tops_burn, eps, steps, iters,
                      use_conv, nb, py, eq_counts, since_burn_out, tops0_out, percent_out, NULL, NULL, NULL);
}

/* ------------------------------------------------------------------ */
/* General (x, y, z) noise: Chain_xyz / _update_chain_fast_xyz          */
/* (mcmc.py:106-114,162-173) and STDC_general_noise(_shortest)          */
/* (decoders.py:325-508).                                               */
/* ------------------------------------------------------------------ */
/* accept iff u < (factors ** (n_new - n_old)).prod(): numpy power per element
 * (libm pow), product accumulated from 1 in index order (numba's array.prod). */
QO_EXPORT void qo_update_chain_fast_xyz(int geom, int L, uint8_t *qm, const double *factors, int64_t iters,
                                        qo_stream *nb)
{
    int n = qo_nsites(geom, L);
    int64_t c0[3], c1[3];
    qo_count_xyz(qm, n, c0);
    for (int64_t t = 0; t < iters; t++) {
        int row, col, op;
        qo_draw_stabilizer(geom, L, nb, &row, &col, &op);
        qo_apply_stabilizer(geom, L, qm, row, col, op);
        qo_count_xyz(qm, n, c1);
        double r = 1.0;
        for (int i = 0; i < 3; i++) r *= pow(factors[i], (double)(c1[i] - c0[i]));
        if (qo_stream_next(nb) < r) { c0[0] = c1[0]; c0[1] = c1[1]; c0[2] = c1[2]; }
        else qo_apply_stabilizer(geom, L, qm, row, col, op);
    }
}

/* weighted length of decoders.py:406: sum of beta_i * n_i over the components with n_i > 0 */
static double xyz_weight(const double *beta, const double *cnt)
{
    double w = 0.0;
    for (int i = 0; i < 3; i++)
        if (cnt[i] > 0) w += beta[i] * cnt[i];
    return w;
}

/* use_xyz != 0: Chain_xyz(p_sampling array) ; else Chain(p_sampling scalar) -- both on the fast path with
 * proposal geometry geom_chain; randomize is False in every branch of the reference (decoders.py:362,375).
 * Outputs: eqdistr (all distinct chains), eqdistr_shortest (chains whose weighted length is np.isclose to the
 * class minimum), both normalised to percent; distinct[n_eq] optional. */
QO_EXPORT void qo_stdc_general_noise(int geom_code, int geom_chain, int L, int n_eq, const uint8_t *qm_init,
                                     const double *p_xyz, int use_xyz, const double *p_sampling_xyz,
                                     double p_sampling, int droplets, int64_t steps, int64_t iters, qo_stream **nb,
                                     double *eqdistr, double *eqdistr_shortest, int64_t *distinct)
{
    int n = qo_nsites(geom_code, L);
    double beta[3], factors[3] = {0, 0, 0};
    for (int i = 0; i < 3; i++) beta[i] = -log((p_xyz[i] / 3) / (1 - p_xyz[i]));
    if (use_xyz) {
        double tot = p_sampling_xyz[0] + p_sampling_xyz[1] + p_sampling_xyz[2];
        for (int i = 0; i < 3; i++) factors[i] = p_sampling_xyz[i] / (1.0 - tot);
    }
    double factor = (p_sampling / 3.0) / (1.0 - p_sampling);
    uint8_t *qm = (uint8_t *)malloc((size_t)n);
    double tot_all = 0, tot_short = 0;
    for (int eq = 0; eq < n_eq; eq++) {
        qo_set *all = qo_set_new(n);
        for (int d = 0; d < droplets; d++) {
            memcpy(qm, qm_init + (size_t)eq * n, (size_t)n);
            for (int64_t s = 0; s < steps; s++) {
                if (use_xyz) qo_update_chain_fast_xyz(geom_chain, L, qm, factors, iters, nb[eq * droplets + d]);
                else qo_update_chain_fast(geom_chain, L, qm, factor, iters, nb[eq * droplets + d], NULL, NULL);
                int nw;
                int64_t e = qo_set_add(all, qm, &nw);
                if (nw) {
                    int64_t c[3];
                    qo_count_xyz(qm, n, c);
                    all->val[3 * e] = (double)c[0]; all->val[3 * e + 1] = (double)c[1]; all->val[3 * e + 2] = (double)c[2];
                }
            }
        }
        double wmin = INFINITY, z = 0, zs = 0;
        for (int64_t e = 0; e < all->cnt; e++) {
            double w = xyz_weight(beta, all->val + 3 * e);
            if (w < wmin) wmin = w;
            z += exp(-w);
        }
        for (int64_t e = 0; e < all->cnt; e++) {
            double w = xyz_weight(beta, all->val + 3 * e);
            if (fabs(w - wmin) <= 1e-8 + 1e-5 * fabs(wmin)) zs += exp(-w); /* np.isclose defaults */
        }
        eqdistr[eq] = z; eqdistr_shortest[eq] = zs;
        tot_all += z; tot_short += zs;
        if (distinct) distinct[eq] = all->cnt;
        qo_set_free(all);
    }
    for (int eq = 0; eq < n_eq; eq++) {
        eqdistr[eq] = eqdistr[eq] / tot_all * 100;
        eqdistr_shortest[eq] = eqdistr_shortest[eq] / tot_short * 100;
    }
    free(qm);
}

/* ------------------------------------------------------------------ */
/* PTDC / PTRC (decoders.py:138-233, 584-742) and                       */
/* PTEQ_alpha_with_shortest (decoders_biasednoise.py:93-172)            */
/* ------------------------------------------------------------------ */
/* One class ladder per class (p_logical = 0), droplets ladders per class, each `steps` Ladder.step(iters)
 * calls; after every step every rung's state is offered to the class's set (PTDC) or to the rung's own set with
 * N(n) / m(n) counters (PTRC).  nb / py: per (class, droplet) streams.
 * mode 0 = PTDC: out = Z_E = sum exp(-beta_error n) over the union, normalised percent (float; the reference
 *                truncates to uint8).
 * mode 1 = PTRC: out = sum over rungs i < Nc-1 of C_mean_i * sum_n m_i(n) exp(n d_beta_i - beta_i n0_i)
 *                (decoders.py:721-739), droplet counters summed (decoders.py:699-718). */
QO_EXPORT void qo_ptxc(int mode, int geom, int L, int n_eq, int Nc, const uint8_t *qm_init, double p_error,
                       double p_sampling, int droplets, int64_t steps, int64_t iters, qo_stream **nb, qo_stream **py,
                       double *out, int64_t *N_out /* mode 1: [n_eq][Nc][n+1] */, int64_t *m_out)
{
    int n = qo_nsites(geom, L);
    double beta_error = -log((p_error / 3) / (1 - p_error));
    double *ladder = (double *)malloc(sizeof(double) * (size_t)Nc);
    double *diff = (double *)calloc((size_t)Nc, sizeof(double));
    /* numpy.linspace(p_sampling, 0.75, Nc) */
    if (Nc == 1) ladder[0] = p_sampling;
    else {
        double step = (0.75 - p_sampling) / (double)(Nc - 1);
        for (int i = 0; i < Nc; i++) { volatile double t = (double)i * step; ladder[i] = t + p_sampling; }
        ladder[Nc - 1] = 0.75;
    }
    for (int i = 0; i + 1 < Nc; i++) diff[i] = (ladder[i] * (1 - ladder[i + 1])) / (ladder[i + 1] * (1 - ladder[i]));
    uint8_t *qm = (uint8_t *)malloc((size_t)n * Nc);
    int32_t *flags = (int32_t *)malloc(sizeof(int32_t) * (size_t)Nc);
    double *n_eff = (double *)calloc((size_t)Nc, sizeof(double));
    int64_t *Nh = (int64_t *)malloc(sizeof(int64_t) * (size_t)Nc * (n + 1));
    int64_t *mh = (int64_t *)malloc(sizeof(int64_t) * (size_t)Nc * (n + 1));
    double total = 0;
    for (int eq = 0; eq < n_eq; eq++) {
        qo_set *all = qo_set_new(n);
        memset(Nh, 0, sizeof(int64_t) * (size_t)Nc * (n + 1));
        memset(mh, 0, sizeof(int64_t) * (size_t)Nc * (n + 1));
        for (int d = 0; d < droplets; d++) {
            qo_set **rung = (qo_set **)malloc(sizeof(qo_set *) * (size_t)Nc);
            for (int i = 0; i < Nc; i++) { rung[i] = qo_set_new(n); memcpy(qm + (size_t)i * n, qm_init + (size_t)eq * n, (size_t)n); flags[i] = 0; }
            flags[Nc - 1] = 1;
            int64_t tops0 = 0;
            for (int64_t s = 0; s < steps; s++) {
                qo_ladder_step(0, geom, L, Nc, qm, ladder, diff, 0.0, 0.0, flags, n_eff, &tops0, iters,
                               nb[eq * droplets + d], py[eq * droplets + d]);
                for (int i = 0; i < Nc; i++) {
                    int nw;
                    const uint8_t *st = qm + (size_t)i * n;
                    if (mode == 0) {
                        int64_t e = qo_set_add(all, st, &nw);
                        if (nw) all->val[3 * e] = qo_count_errors(st, n);
                    } else {
                        int64_t e = qo_set_add(rung[i], st, &nw);
                        if (nw) { rung[i]->val[3 * e] = qo_count_errors(st, n); Nh[(size_t)i * (n + 1) + (int)rung[i]->val[3 * e]]++; }
                        mh[(size_t)i * (n + 1) + (int)rung[i]->val[3 * e]]++;
                    }
                }
            }
            for (int i = 0; i < Nc; i++) qo_set_free(rung[i]);
            free(rung);
        }
        double z = 0;
        if (mode == 0) {
            for (int64_t e = 0; e < all->cnt; e++) z += exp(-beta_error * all->val[3 * e]);
        } else {
            for (int i = 0; i < Nc - 1; i++) {
                double beta_i = -log((ladder[i] / 3) / (1 - ladder[i])), d_beta = beta_i - beta_error;
                const int64_t *N = Nh + (size_t)i * (n + 1), *m = mh + (size_t)i * (n + 1);
                int l0 = -1, l1 = -1;
                for (int l = 0; l <= n; l++)
                    if (m[l]) { if (l0 < 0) l0 = l; else if (l1 < 0) l1 = l; }
                double c_mean = (double)N[l0] / (double)m[l0] * exp(-beta_i * 0.0);
                if (l1 >= 0) c_mean = (c_mean + (double)N[l1] / (double)m[l1] * exp(-beta_i * (double)(l1 - l0))) / 2.0;
                double sum = 0;
                for (int l = 0; l <= n; l++)
                    if (m[l]) sum += (double)m[l] * exp((double)l * d_beta - beta_i * (double)l0);
                z += c_mean * sum;
            }
            if (N_out) memcpy(N_out + (size_t)eq * Nc * (n + 1), Nh, sizeof(int64_t) * (size_t)Nc * (n + 1));
            if (m_out) memcpy(m_out + (size_t)eq * Nc * (n + 1), mh, sizeof(int64_t) * (size_t)Nc * (n + 1));
        }
        out[eq] = z;
        total += z;
        qo_set_free(all);
    }
    for (int eq = 0; eq < n_eq; eq++) out[eq] = out[eq] / total * 100;
    free(ladder); free(diff); free(qm); free(flags); free(n_eff); free(Nh); free(mh);
}

/* PTDC with the early stop of PTDC_droplet (decoders.py:138-165): every droplet keeps its own `samples` dictionary over
 * all rungs; a chain new to the droplet whose length is <= the shortest seen so far moves `stop` to step * conv_mult, and the
 * droplet ends once step >= stop and step * 100 >= steps.  The class's Z_E runs over the union of the droplets' samples
 * (decoders.py:221-231).  steps_done [n_eq * droplets]: Ladder.step calls made; N_out [n_eq][n + 1]: distinct chains per length. */
QO_EXPORT void qo_ptdc_conv(int geom, int L, int n_eq, int Nc, const uint8_t *qm_init, double p_error, double p_sampling,
                            int droplets, int64_t steps, int64_t iters, double conv_mult, qo_stream **nb, qo_stream **py,
                            double *out, int64_t *steps_done, int64_t *N_out)
{
    int n = qo_nsites(geom, L);
    double beta_error = -log((p_error / 3) / (1 - p_error));
    double *ladder = (double *)malloc(sizeof(double) * (size_t)Nc);
    double *diff = (double *)calloc((size_t)Nc, sizeof(double));
    if (Nc == 1) ladder[0] = p_sampling;
    else {
        double step = (0.75 - p_sampling) / (double)(Nc - 1);
        for (int i = 0; i < Nc; i++) { volatile double t = (double)i * step; ladder[i] = t + p_sampling; }
        ladder[Nc - 1] = 0.75;
    }
    for (int i = 0; i + 1 < Nc; i++) diff[i] = (ladder[i] * (1 - ladder[i + 1])) / (ladder[i + 1] * (1 - ladder[i]));
    uint8_t *qm = (uint8_t *)malloc((size_t)n * Nc);
    int32_t *flags = (int32_t *)malloc(sizeof(int32_t) * (size_t)Nc);
    double *n_eff = (double *)calloc((size_t)Nc, sizeof(double));
    double total = 0;
    for (int eq = 0; eq < n_eq; eq++) {
        qo_set *all = qo_set_new(n);
        for (int d = 0; d < droplets; d++) {
            qo_set *mine = qo_set_new(n);
            for (int i = 0; i < Nc; i++) { memcpy(qm + (size_t)i * n, qm_init + (size_t)eq * n, (size_t)n); flags[i] = 0; }
            flags[Nc - 1] = 1;
            int64_t tops0 = 0, s = 0;
            int shortest = 2 * L * L;
            double stop = (double)steps;
            for (; s < steps; s++) {
                qo_ladder_step(0, geom, L, Nc, qm, ladder, diff, 0.0, 0.0, flags, n_eff, &tops0, iters,
                               nb[eq * droplets + d], py[eq * droplets + d]);
                for (int i = 0; i < Nc; i++) {
                    int nw, nw2;
                    const uint8_t *st = qm + (size_t)i * n;
                    qo_set_add(mine, st, &nw);
                    if (nw) {
                        int length = qo_count_errors(st, n);
                        int64_t e = qo_set_add(all, st, &nw2);
                        if (nw2) all->val[3 * e] = length;
                        if (conv_mult != 0 && length <= shortest) { shortest = length; stop = (double)s * conv_mult; }
                    }
                }
                if (conv_mult != 0 && (double)s >= stop && s * 100 >= steps) { s++; break; }
            }
            if (steps_done) steps_done[eq * droplets + d] = s;
            qo_set_free(mine);
        }
        double z = 0;
        for (int64_t e = 0; e < all->cnt; e++) {
            z += exp(-beta_error * all->val[3 * e]);
            if (N_out) N_out[(size_t)eq * (n + 1) + (int)all->val[3 * e]]++;
        }
        out[eq] = z;
        total += z;
        qo_set_free(all);
    }
    for (int eq = 0; eq < n_eq; eq++) out[eq] = out[eq] / total * 100;
    free(ladder); free(diff); free(qm); free(flags); free(n_eff);
}

/* ------------------------------------------------------------------ */
/* CPU baseline driver: a range of syndromes, single-threaded; the      */
/* Python wrapper fans ranges out over host threads (ctypes drops the   */
/* GIL).  One independent NB/NP stream pair per (syndrome, class,       */
/* droplet), seeded from (seed, global indices).  bench.py times this.  */
/* ------------------------------------------------------------------ */
/* One (syndrome, class) of qo_stdc_batch: the unnormalised Z_E of class eq -- the CPU baseline's unit of parallel
 * work when a step holds fewer syndromes than the host has threads.  Same streams as qo_stdc_batch. */
QO_EXPORT double qo_stdc_class(int geom_code, int geom_chain, int L, const uint8_t *qm /* [n_sites] */, int eq,
                               double p_error, double p_sampling, int droplets, int64_t steps, int64_t iters,
                               uint32_t seed, int64_t s_index)
{
    int n = qo_nsites(geom_code, L), n_eq = qo_neq(geom_code);
    double factor = (p_sampling / 3.0) / (1.0 - p_sampling), beta = -log((p_error / 3) / (1 - p_error));
    uint8_t *init = (uint8_t *)malloc((size_t)n), *cur = (uint8_t *)malloc((size_t)n);
    memcpy(init, qm, (size_t)n);
    qo_to_class(geom_code, L, init, eq);
    qo_set *all = qo_set_new(n);
    for (int d = 0; d < droplets; d++) {
        uint32_t id = (uint32_t)((s_index * n_eq + eq) * droplets + d);
        qo_stream *nb = qo_stream_mt(seed ^ (2u * id + 1u) * 2654435761u);
        qo_stream *np_ = qo_stream_mt(seed ^ (2u * id + 2u) * 2246822519u);
        memcpy(cur, init, (size_t)n);
        qo_set *s = qo_stdc_droplet(geom_code, geom_chain, L, cur, factor, steps, iters, 1, 0.0, nb, np_, NULL);
        qo_set_update(all, s);
        qo_set_free(s);
        qo_stream_free(nb); qo_stream_free(np_);
    }
    double z = 0;
    for (int64_t e = 0; e < all->cnt; e++) z += exp(-beta * all->val[3 * e]);
    qo_set_free(all);
    free(init); free(cur);
    return z;
}

QO_EXPORT void qo_stdc_batch(int geom_code, int geom_chain, int L, int64_t S, const uint8_t *qm /* [S][n_sites] */,
                             double p_error, double p_sampling, int droplets, int64_t steps, int64_t iters,
                             uint32_t seed, int64_t s_offset, double *eqdistr /* [S][n_eq] */)
{
    int n = qo_nsites(geom_code, L), n_eq = qo_neq(geom_code);
    for (int64_t s = 0; s < S; s++) {
        uint8_t *init = (uint8_t *)malloc((size_t)n * n_eq);
        qo_stream **nb = (qo_stream **)malloc(sizeof(qo_stream *) * (size_t)n_eq * droplets);
        qo_stream **np_ = (qo_stream **)malloc(sizeof(qo_stream *) * (size_t)n_eq * droplets);
        for (int eq = 0; eq < n_eq; eq++) {
            memcpy(init + (size_t)eq * n, qm + (size_t)s * n, (size_t)n);
            qo_to_class(geom_code, L, init + (size_t)eq * n, eq);
            for (int d = 0; d < droplets; d++) {
                uint32_t id = (uint32_t)(((s + s_offset) * n_eq + eq) * droplets + d);
                nb[eq * droplets + d] = qo_stream_mt(seed ^ (2u * id + 1u) * 2654435761u);
                np_[eq * droplets + d] = qo_stream_mt(seed ^ (2u * id + 2u) * 2246822519u);
            }
        }
        qo_stdc(geom_code, geom_chain, L, n_eq, init, p_error, p_sampling, droplets, steps, iters, 1, 0.0, nb, np_,
                eqdistr + (size_t)s * n_eq, NULL, NULL);
        for (int i = 0; i < n_eq * droplets; i++) { qo_stream_free(nb[i]); qo_stream_free(np_[i]); }
        free(init); free(nb); free(np_);
    }
}
