// qecmc_lattice.h -- 2-bit-packed Pauli lattices and stabilizer geometry.
//
// Pauli per qubit in two bits (0=I, 1=X, 2=Y, 3=Z; composition is XOR, as in the
// reference's uint8 lattices).  One "row word" holds one lattice row:
// column c lives in bits [2c, 2c+2).  Toric/planar: word index = layer*L + row
// (2L words); rotated/XZZX: word index = row (L words).  Word type is uint32_t for
// L <= 16 and uint64_t for L <= 32.
//
// Everything here is __host__ __device__ so that tests/ can compile it with g++
// and check it against the oracle without a GPU; the product only ever calls it
// from CUDA kernels.
//
// Geometry follows the reference (cited per function); nothing here is derived
// from its source text -- stabilizers are expressed as XOR masks on row words.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define QHD __host__ __device__ __forceinline__
#else
#define QHD inline
#endif

namespace qecmc {

enum { TORIC = 0, PLANAR = 1, ROTATED = 2, XZZX = 3 };

QHD int popc(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
QHD int popc(uint64_t x)
{
#if defined(__CUDA_ARCH__)
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}

template <typename W> struct WordTraits;
template <> struct WordTraits<uint32_t> { static constexpr uint32_t LO = 0x55555555u; };
template <> struct WordTraits<uint64_t> { static constexpr uint64_t LO = 0x5555555555555555ull; };

// low bit of every non-identity field
template <typename W> QHD W nzmap(W w) { return (W)((w | (w >> 1)) & WordTraits<W>::LO); }
// number of non-identity qubits in a row word (count_errors restricted to a row)
template <typename W> QHD int weight(W w) { return popc(nzmap(w)); }
// fields equal to X (01), Y (10), Z (11)
template <typename W> QHD W xmap(W w) { return (W)(w & ~(w >> 1) & WordTraits<W>::LO); }
template <typename W> QHD W ymap(W w) { return (W)(~w & (w >> 1) & WordTraits<W>::LO); }
template <typename W> QHD W zmap(W w) { return (W)(w & (w >> 1) & WordTraits<W>::LO); }

template <typename W> QHD W fld(int pauli, int col) { return (W)((W)pauli << (2 * col)); }

struct Geo {
    int geom, L, nw, nstab, nsites, neq, layers;
    int nfull;           // toric: L*L, planar: L*(L-1), rotated/xzzx: (L-1)^2
    uint32_t magicL;     // ceil(2^16 / L)      -> n / L     for n < 2048
    uint32_t magicLm1;   // ceil(2^16 / (L-1))  -> n / (L-1) for n < 2048
};

QHD Geo make_geo(int geom, int L)
{
    Geo g;
    g.geom = geom;
    g.L = L;
    g.layers = (geom == TORIC || geom == PLANAR) ? 2 : 1;
    g.nw = g.layers * L;
    g.nsites = g.layers * L * L;
    g.neq = geom == TORIC ? 16 : 4;
    if (geom == TORIC) { g.nfull = L * L; g.nstab = 2 * L * L; }
    else if (geom == PLANAR) { g.nfull = L * (L - 1); g.nstab = 2 * L * (L - 1); }
    else { g.nfull = (L - 1) * (L - 1); g.nstab = L * L - 1; }
    g.magicL = (65536u + (uint32_t)L - 1u) / (uint32_t)L;
    g.magicLm1 = L > 1 ? (65536u + (uint32_t)L - 2u) / (uint32_t)(L - 1) : 65536u;
    return g;
}

// exact n / d for n < 2048, d <= 32 with magic = ceil(2^16/d)
QHD int fastdiv(int n, uint32_t magic) { return (int)(((uint32_t)n * magic) >> 16); }

// Up to three touched row words per stabilizer (two for the single-layer codes).
template <typename W> struct Upd {
    int w[3];
    W m[3];
};
template <int GEOM> struct NumUpd { static constexpr int value = (GEOM == TORIC || GEOM == PLANAR) ? 3 : 2; };

// Canonical stabilizer index <-> the reference's (row, col, operator) triple.
// All stabilizers of a code are equiprobable under the reference's proposal
// (toric_model.py:287-296; planar_model.py:342-352; rotated_surface_model.py:395-408
// and xzzx_model.py:439-452, where P(full)=1-2/(L+1) over (L-1)^2 sites and
// P(half)=2/(L+1) over 2(L-1) sites both give 1/(L^2-1) per stabilizer).
template <int GEOM> QHD void idx_to_rco(const Geo &g, int idx, int &row, int &col, int &op)
{
    if (GEOM == TORIC) {
        op = idx < g.nfull ? 1 : 3;
        int rem = idx < g.nfull ? idx : idx - g.nfull;
        row = fastdiv(rem, g.magicL);
        col = rem - row * g.L;
    } else if (GEOM == PLANAR) {
        if (idx < g.nfull) { op = 1; row = fastdiv(idx, g.magicL); col = idx - row * g.L; }
        else { int rem = idx - g.nfull; op = 3; row = fastdiv(rem, g.magicLm1); col = rem - row * (g.L - 1); }
    } else {
        if (idx < g.nfull) { op = 1; row = fastdiv(idx, g.magicLm1); col = idx - row * (g.L - 1); }
        else { int rem = idx - g.nfull; op = 3; row = rem >> 2; col = rem & 3; }
    }
}

template <int GEOM> QHD int rco_to_idx(const Geo &g, int row, int col, int op)
{
    if (GEOM == TORIC) return (op == 1 ? 0 : g.nfull) + row * g.L + col;
    if (GEOM == PLANAR) return op == 1 ? row * g.L + col : g.nfull + row * (g.L - 1) + col;
    return op == 1 ? row * (g.L - 1) + col : g.nfull + row * 4 + col;
}

// Stabilizer (row, col, op) -> XOR masks on row words.
// toric_model.py:256-284, planar_model.py:291-339,
// rotated_surface_model.py:349-392, xzzx_model.py:360-436.
template <int GEOM, typename W> QHD void decode(const Geo &g, int row, int col, int op, Upd<W> &u)
{
    const int L = g.L;
    if (GEOM == TORIC) {
        if (op == 1) {  // X on [1,r,c] [1,r,c-1] [0,r,c] [0,r-1,c]
            int cm = col == 0 ? L - 1 : col - 1;
            u.w[0] = L + row; u.m[0] = (W)(fld<W>(1, col) | fld<W>(1, cm));
            u.w[1] = row; u.m[1] = fld<W>(1, col);
            u.w[2] = row == 0 ? L - 1 : row - 1; u.m[2] = fld<W>(1, col);
        } else {        // Z on [1,r,c] [0,r,c] [0,r,c+1] [1,r+1,c]
            int cp = col == L - 1 ? 0 : col + 1;
            u.w[0] = row; u.m[0] = (W)(fld<W>(3, col) | fld<W>(3, cp));
            u.w[1] = L + row; u.m[1] = fld<W>(3, col);
            u.w[2] = L + (row == L - 1 ? 0 : row + 1); u.m[2] = fld<W>(3, col);
        }
    } else if (GEOM == PLANAR) {
        if (op == 1) {  // X on [0,r,c] [0,r+1,c] (+[1,r,c] if c<L-1) (+[1,r,c-1] if c>0)
            W m = 0;
            if (col < L - 1) m |= fld<W>(1, col);
            if (col > 0) m |= fld<W>(1, col - 1);
            u.w[0] = L + row; u.m[0] = m;
            u.w[1] = row; u.m[1] = fld<W>(1, col);
            u.w[2] = row + 1; u.m[2] = fld<W>(1, col);
        } else {        // Z on [0,r,c] [0,r,c+1] (+[1,r,c] if r<L-1) (+[1,r-1,c] if r>0)
            u.w[0] = row; u.m[0] = (W)(fld<W>(3, col) | fld<W>(3, col + 1));
            u.w[1] = L + row; u.m[1] = row < L - 1 ? fld<W>(3, col) : (W)0;
            u.w[2] = row > 0 ? L + row - 1 : L + 1; u.m[2] = row > 0 ? fld<W>(3, col) : (W)0;  // row 0: spare word != w[1]
        }
    } else if (GEOM == ROTATED) {
        if (op == 1) {  // full plaquette, X if (r+c) even else Z
            int p = ((row + col) & 1) ? 3 : 1;
            W m = (W)(fld<W>(p, col) | fld<W>(p, col + 1));
            u.w[0] = row; u.m[0] = m;
            u.w[1] = row + 1; u.m[1] = m;
        } else {        // half plaquette: row = k, col = side
            int k = row;
            if (col == 0) { u.w[0] = 0; u.m[0] = (W)(fld<W>(1, 2 * k + 1) | fld<W>(1, 2 * k + 2)); u.w[1] = 1; u.m[1] = 0; }
            else if (col == 1) { u.w[0] = 2 * k + 1; u.m[0] = fld<W>(3, L - 1); u.w[1] = 2 * k + 2; u.m[1] = fld<W>(3, L - 1); }
            else if (col == 2) { u.w[0] = L - 1; u.m[0] = (W)(fld<W>(1, 2 * k) | fld<W>(1, 2 * k + 1)); u.w[1] = 0; u.m[1] = 0; }
            else { u.w[0] = 2 * k; u.m[0] = fld<W>(3, 0); u.w[1] = 2 * k + 1; u.m[1] = fld<W>(3, 0); }
        }
        u.w[2] = 0; u.m[2] = 0;
    } else {  // XZZX
        if (op == 1) {  // [r,c]^=X [r+1,c]^=Z [r,c+1]^=Z [r+1,c+1]^=X
            u.w[0] = row; u.m[0] = (W)(fld<W>(1, col) | fld<W>(3, col + 1));
            u.w[1] = row + 1; u.m[1] = (W)(fld<W>(3, col) | fld<W>(1, col + 1));
        } else {
            int k = row;
            if (col == 0) { u.w[0] = 0; u.m[0] = (W)(fld<W>(3, 2 * k + 1) | fld<W>(1, 2 * k + 2)); u.w[1] = 1; u.m[1] = 0; }
            else if (col == 1) { u.w[0] = 2 * k + 1; u.m[0] = fld<W>(1, L - 1); u.w[1] = 2 * k + 2; u.m[1] = fld<W>(3, L - 1); }
            else if (col == 2) { u.w[0] = L - 1; u.m[0] = (W)(fld<W>(1, 2 * k) | fld<W>(3, 2 * k + 1)); u.w[1] = 0; u.m[1] = 0; }
            else { u.w[0] = 2 * k; u.m[0] = fld<W>(3, 0); u.w[1] = 2 * k + 1; u.m[1] = fld<W>(1, 0); }
        }
        u.w[2] = 0; u.m[2] = 0;
    }
}

// ---------------------------------------------------------------------------
// Linear (GF(2)) fingerprint: h(state) = XOR of K(word, bit) over set bits, so
// h(state ^ mask) = h(state) ^ h(mask) and an accepted stabilizer costs one XOR
// with a per-stabilizer constant.  K is a splitmix64 stream of the bit position.
// Stands in for hash(qubit_matrix.tobytes()) (decoders.py:251); any 64-bit
// fingerprint is statistically equivalent.
// ---------------------------------------------------------------------------
QHD uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
QHD uint64_t bitkey(uint64_t seed, int word, int bit) { return mix64(seed ^ (uint64_t)(word * 64 + bit) * 0xD6E8FEB86659FD93ull); }

template <typename W> QHD uint64_t word_hash(uint64_t seed, int word, W bits)
{
    uint64_t h = 0;
    for (int b = 0; b < (int)(8 * sizeof(W)); b++)
        if ((bits >> b) & 1) h ^= bitkey(seed, word, b);
    return h;
}

template <int GEOM, typename W> QHD uint64_t stab_hash(const Geo &g, int idx, uint64_t seed)
{
    int row, col, op;
    idx_to_rco<GEOM>(g, idx, row, col, op);
    Upd<W> u;
    decode<GEOM, W>(g, row, col, op, u);
    uint64_t h = 0;
    for (int i = 0; i < NumUpd<GEOM>::value; i++) h ^= word_hash<W>(seed, u.w[i], u.m[i]);
    return h;
}

// ---------------------------------------------------------------------------
// Whole-lattice operations through an accessor A with  W get(int w) / void set(int w, W).
// ---------------------------------------------------------------------------
template <typename W, typename A> QHD int lat_weight(const Geo &g, const A &a)
{
    int n = 0;
    for (int w = 0; w < g.nw; w++) n += weight<W>(a.get(w));
    return n;
}

template <typename W, typename A> QHD uint64_t lat_hash(const Geo &g, const A &a, uint64_t seed)
{
    uint64_t h = 0;
    for (int w = 0; w < g.nw; w++) h ^= word_hash<W>(seed, w, a.get(w));
    return h;
}

template <typename W, typename A> QHD void lat_count_xyz(const Geo &g, const A &a, int &nx, int &ny, int &nz)
{
    nx = ny = nz = 0;
    for (int w = 0; w < g.nw; w++) {
        W v = a.get(w);
        nx += popc(xmap(v)); ny += popc(ymap(v)); nz += popc(zmap(v));
    }
}

// apply an update (load every touched word, then store every word -- the order the
// kernels use, so the word indices of an update must be pairwise distinct);
// returns the weight change
template <int GEOM, typename W, typename A> QHD int lat_apply(A &a, const Upd<W> &u)
{
    constexpr int NU = NumUpd<GEOM>::value;
    int d = 0;
    W n[NU];
    for (int i = 0; i < NU; i++) {
        W o = a.get(u.w[i]);
        n[i] = (W)(o ^ u.m[i]);
        d += weight<W>(n[i]) - weight<W>(o);
    }
    for (int i = 0; i < NU; i++) a.set(u.w[i], n[i]);
    return d;
}

// _define_equivalence_class: toric_model.py:317-351, planar_model.py:379-390,
// rotated_surface_model.py:411-420, xzzx_model.py:455-486.
template <int GEOM, typename W, typename A> QHD int lat_class(const Geo &g, const A &a)
{
    const int L = g.L;
    const W LO = WordTraits<W>::LO;
    if (GEOM == TORIC) {
        int par[4] = {0, 0, 0, 0};
        for (int l = 0; l < 2; l++)
            for (int r = 0; r < L; r++) {
                W v = a.get(l * L + r);
                par[2 * l] ^= popc((W)((v ^ (v >> 1)) & LO)) & 1;  // X or Y
                par[2 * l + 1] ^= popc((W)((v >> 1) & LO)) & 1;    // Z or Y
            }
        return par[0] + 2 * par[1] + 4 * par[2] + 8 * par[3];
    }
    if (GEOM == PLANAR || GEOM == ROTATED) {
        // planar: X|Y parity in column 0 of layer 0, Z|Y parity in row 0 of layer 0
        // rotated: X|Y parity in row 0, Z|Y parity in column 0
        int colpar = 0;
        for (int r = 0; r < L; r++) {
            W v = a.get(r);
            colpar ^= (GEOM == PLANAR) ? (int)((v ^ (v >> 1)) & 1) : (int)((v >> 1) & 1);
        }
        W v0 = a.get(0);
        int rowpar = (GEOM == PLANAR) ? (popc((W)((v0 >> 1) & LO)) & 1) : (popc((W)((v0 ^ (v0 >> 1)) & LO)) & 1);
        return (GEOM == PLANAR) ? colpar + 2 * rowpar : rowpar + 2 * colpar;
    }
    // XZZX: row 0 counts Y always, X at even col, Z at odd col; column 0 counts Y always,
    // Z at even row, X at odd row.
    int x = 0, z = 0;
    W v0 = a.get(0);
    for (int c = 0; c < L; c++) {
        int q = (int)((v0 >> (2 * c)) & 3);
        x ^= (q == 2) || ((c & 1) ? q == 3 : q == 1);
    }
    for (int r = 0; r < L; r++) {
        int q = (int)(a.get(r) & 3);
        z ^= (q == 2) || ((r & 1) ? q == 1 : q == 3);
    }
    return x ? (z ? 2 : 1) : (z ? 3 : 0);
}

template <typename W> QHD W rowmask(int pauli, int L)
{
    // the Pauli in every one of the L fields: 0b0101.. times the Pauli, cut to 2L bits
    const W all = (W)((W)(~(W)0 / 3) * (W)pauli);
    return 2 * L >= (int)(8 * sizeof(W)) ? all : (W)(all & (W)(((W)1 << (2 * L)) - 1));
}

template <typename W, typename A> QHD int xor_word(A &a, int w, W m)
{
    W o = a.get(w), n = (W)(o ^ m);
    a.set(w, n);
    return weight<W>(n) - weight<W>(o);
}

// _apply_logical: toric_model.py:179-225, planar_model.py:234-268,
// rotated_surface_model.py:251-282, xzzx_model.py:279-313.  Returns the weight change.
template <int GEOM, typename W, typename A>
QHD int lat_apply_logical(const Geo &g, A &a, int op, int layer, int X_pos, int Z_pos)
{
    const int L = g.L;
    int d = 0;
    if (op == 0) return 0;
    if (GEOM == TORIC) {
        bool do_X = (op == 1 || op == 2), do_Z = (op == 3 || op == 2);
        if (layer == 0) {
            // X along row X_pos and Z along column Z_pos may share one qubit: apply sequentially
            if (do_X) d += xor_word<W>(a, X_pos, rowmask<W>(1, L));
            if (do_Z) for (int i = 0; i < L; i++) d += xor_word<W>(a, i, fld<W>(3, Z_pos));
        } else {  // layer 1 is addressed transposed: X on [1,i,X_pos], Z on [1,Z_pos,i]
            if (do_X) for (int i = 0; i < L; i++) d += xor_word<W>(a, L + i, fld<W>(1, X_pos));
            if (do_Z) d += xor_word<W>(a, L + Z_pos, rowmask<W>(3, L));
        }
    } else if (GEOM == PLANAR) {
        bool do_X = (op == 1 || op == 3), do_Z = (op == 2 || op == 3);
        if (do_X) d += xor_word<W>(a, X_pos, rowmask<W>(1, L));
        if (do_Z) for (int i = 0; i < L; i++) d += xor_word<W>(a, i, fld<W>(3, Z_pos));
    } else if (GEOM == ROTATED) {
        bool do_X = (op == 1 || op == 3), do_Z = (op == 2 || op == 3);
        if (do_X) for (int i = 0; i < L; i++) d += xor_word<W>(a, i, fld<W>(1, X_pos));
        if (do_Z) d += xor_word<W>(a, Z_pos, rowmask<W>(3, L));
    } else {
        bool do_X = (op == 1 || op == 2), do_Z = (op == 3 || op == 2);
        if (do_X) for (int i = 0; i < L; i++) d += xor_word<W>(a, i, fld<W>(1, L - 1 - i));
        if (do_Z) for (int i = 0; i < L; i++) d += xor_word<W>(a, i, fld<W>(3, i));
    }
    return d;
}

// The same operators as XOR masks per row word: mask of word w for the operator(s) of _apply_random_logical
// (op0 / X0 / Z0 on layer 0 -- the only layer of the one-layer codes --, op1 / X1 / Z1 on the toric code's layer 1).
// Applying the masks to every word equals lat_apply_logical layer by layer (XOR commutes; X ^ Z = Y on a shared qubit).
template <int GEOM, typename W> QHD W logical_mask(const Geo &g, int w, int op0, int op1, int X0, int Z0, int X1, int Z1)
{
    const int L = g.L;
    W m = 0;
    if (GEOM == TORIC) {
        if (w < L) {
            if ((op0 == 1 || op0 == 2) && w == X0) m ^= rowmask<W>(1, L);
            if (op0 == 3 || op0 == 2) m ^= fld<W>(3, Z0);
        } else {
            const int i = w - L;
            if (op1 == 1 || op1 == 2) m ^= fld<W>(1, X1);
            if ((op1 == 3 || op1 == 2) && i == Z1) m ^= rowmask<W>(3, L);
        }
    } else if (GEOM == PLANAR) {
        if (w < L) {
            if ((op0 == 1 || op0 == 3) && w == X0) m ^= rowmask<W>(1, L);
            if (op0 == 2 || op0 == 3) m ^= fld<W>(3, Z0);
        }
    } else if (GEOM == ROTATED) {
        if (op0 == 1 || op0 == 3) m ^= fld<W>(1, X0);
        if ((op0 == 2 || op0 == 3) && w == Z0) m ^= rowmask<W>(3, L);
    } else {
        if (op0 == 1 || op0 == 2) m ^= fld<W>(1, L - 1 - w);
        if (op0 == 3 || op0 == 2) m ^= fld<W>(3, w);
    }
    return m;
}

// Move into class eq keeping the syndrome.  Toric: _to_class, toric_model.py:354-377;
// others: apply_logical(define_equivalence_class() ^ eq), decoders.py:556-560.
template <int GEOM, typename W, typename A> QHD void lat_to_class(const Geo &g, A &a, int eq)
{
    int diff = eq ^ lat_class<GEOM, W>(g, a);
    if (GEOM == TORIC) {
        int ops = diff ^ ((diff & 0xA) >> 1);
        lat_apply_logical<GEOM, W>(g, a, ops & 3, 0, 0, 0);
        lat_apply_logical<GEOM, W>(g, a, ops >> 2, 1, 0, 0);
    } else {
        lat_apply_logical<GEOM, W>(g, a, diff, 0, 0, 0);
    }
}

// "Rain" legality: planar_model.py:363-365 removes x-operators in the last row and
// z-operators in the last column; toric has none.  (o, r, c) index np.random.rand(2,L,L);
// o == 0 means operator 3, o == 1 operator 1 (toric_model.py:309-312).
template <int GEOM> QHD bool rain_legal(const Geo &g, int o, int r, int c)
{
    if (GEOM == PLANAR) return !((o == 1 && r == g.L - 1) || (o == 0 && c == g.L - 1));
    return true;
}

// bytes (reference layout, C order) <-> row words
template <typename W> QHD W pack_row(const uint8_t *row, int L)
{
    W w = 0;
    for (int c = 0; c < L; c++) w |= (W)((W)(row[c] & 3) << (2 * c));
    return w;
}
template <typename W> QHD void unpack_row(W w, uint8_t *row, int L)
{
    for (int c = 0; c < L; c++) row[c] = (uint8_t)((w >> (2 * c)) & 3);
}

}  // namespace qecmc
