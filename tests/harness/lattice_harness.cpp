// Host-side build of csrc/qecmc_lattice.h for CPU unit tests (tests/test_lattice_host.py).
// Test infrastructure only: lets the packed-lattice geometry be checked against the
// oracle without a GPU.  The product never links this file.
#include <string.h>
#include <vector>
#include "../../mcmc-qec-toric-rl_b200/csrc/qecmc_lattice.h"
using namespace qecmc;

template <typename W> struct VecAcc {
    W *p;
    W get(int w) const { return p[w]; }
    void set(int w, W v) { p[w] = v; }
};

template <typename W> static void pack(const Geo &g, const uint8_t *qm, std::vector<W> &v)
{
    v.resize(g.nw);
    for (int w = 0; w < g.nw; w++) v[w] = pack_row<W>(qm + (size_t)w * g.L, g.L);
}
template <typename W> static void unpack(const Geo &g, const std::vector<W> &v, uint8_t *qm)
{
    for (int w = 0; w < g.nw; w++) unpack_row<W>(v[w], qm + (size_t)w * g.L, g.L);
}

// what: 0 apply stabilizer (a=row,b=col,c=op) by rco; 1 apply stabilizer by canonical idx (a=idx);
//       2 apply logical (a=op,b=layer,c=X_pos,d=Z_pos); 3 to_class (a=eq); 4 class; 5 weight;
//       6 idx->rco->idx roundtrip check (a=idx) returns idx'; 7 hash of state (lo 31 bits) ;
//       8 hash linearity check for stabilizer idx a: returns 1 if h(s^m)==h(s)^stab_hash
template <int GEOM, typename W> static int run(int L, uint8_t *qm, int what, int a, int b, int c, int d)
{
    Geo g = make_geo(GEOM, L);
    std::vector<W> v;
    pack<W>(g, qm, v);
    VecAcc<W> acc{v.data()};
    int ret = 0;
    Upd<W> u;
    const uint64_t seed = 0x1234abcdull;
    switch (what) {
    case 0: decode<GEOM, W>(g, a, b, c, u); ret = lat_apply<GEOM, W>(acc, u); break;
    case 1: { int r, cc, op; idx_to_rco<GEOM>(g, a, r, cc, op); decode<GEOM, W>(g, r, cc, op, u); ret = lat_apply<GEOM, W>(acc, u); break; }
    case 2: ret = lat_apply_logical<GEOM, W>(g, acc, a, b, c, d); break;
    case 3: lat_to_class<GEOM, W>(g, acc, a); break;
    case 4: ret = lat_class<GEOM, W>(g, acc); break;
    case 5: ret = lat_weight<W>(g, acc); break;
    case 6: { int r, cc, op; idx_to_rco<GEOM>(g, a, r, cc, op); ret = rco_to_idx<GEOM>(g, r, cc, op); break; }
    case 7: ret = (int)(lat_hash<W>(g, acc, seed) & 0x7fffffff); break;
    case 8: {
        uint64_t h0 = lat_hash<W>(g, acc, seed);
        int r, cc, op; idx_to_rco<GEOM>(g, a, r, cc, op); decode<GEOM, W>(g, r, cc, op, u); lat_apply<GEOM, W>(acc, u);
        ret = (lat_hash<W>(g, acc, seed) == (h0 ^ stab_hash<GEOM, W>(g, a, seed)));
        break; }
    }
    unpack<W>(g, v, qm);
    return ret;
}

extern "C" int lh_nstab(int geom, int L) { return make_geo(geom, L).nstab; }

extern "C" int lh_run(int geom, int L, int wide, uint8_t *qm, int what, int a, int b, int c, int d)
{
#define CASE(G) case G: return wide ? run<G, uint64_t>(L, qm, what, a, b, c, d) : run<G, uint32_t>(L, qm, what, a, b, c, d);
    switch (geom) { CASE(TORIC) CASE(PLANAR) CASE(ROTATED) CASE(XZZX) }
    return -999;
}
