// tile_metropolis.cuh -- production Metropolis: shared-memory boxes with sublattice-parallel swaps.
//
// Replaces the k-loop of metropolis_simulated_annealing (src/metropolis.F90:348-354) and
// monte_carlo_step_lattice / _nbr (:751-891) for lattices large enough to be decomposed.
//
// Scheme (DESIGN.md "Production Metropolis"):
//  * PHASE: the periodic lattice is cut into boxes of (Bx,By,Bz) doubled-grid units at a random
//    even origin; one CTA loads one box (sites only, 1 B/site) into shared memory.  Sites closer
//    than `m` (>= interaction reach) to a box face are frozen during the phase, so every
//    neighbour of an active site is inside the same box: no halo, no inter-CTA traffic.
//  * STEP: a CTA-uniform random draw picks a residue class o (mod P) and an allowed
//    displacement class d.  Thread (i,j,k) proposes the swap of site1 = m + o + P*(i,j,k) with
//    site2 = m + (o+d mod P) + P*((i,j,k)+s mod A).  P and the set D of allowed d are chosen
//    on the host such that NO neighbour vector v of the Hamiltonian satisfies v = 0, +-d (mod P):
//    all sites touched in one step are mutually non-interacting, hence the M = Ax*Ay*Az
//    simultaneous Metropolis decisions are exactly equivalent to M sequential ones (detailed
//    balance holds move by move; the proposal is symmetric and configuration-independent).
//  * dE uses the reference's f64 association (brw_common.cuh) => each decision is the one the
//    reference would take for the same pair, same configuration, same uniform.
//  * RNG: Philox4x32-10, counter = (trial slot, step/4, box, phase), key = seed.
#pragma once
#include <vector>
#include <algorithm>
#include "brawl_common.cuh"

#define BRW_MAX_MODES 3
struct BrwBoxMode {            // one orientation of the (rectangular) period, see brw_choose_modes
  int P[3];                    // period per axis (doubled-grid units, even)
  int A[3], M;                 // coarse cells per axis inside the active region, M = A0*A1*A2 trials per step
  int cls0, n_classes;         // slice of the residue-class table
  int d0, n_disp;              // slice of the displacement-class table
};
struct BrwBoxParams {          // POD kernel parameter
  double guard;                // screening guard band on dE (Ry); see brw_box_metropolis_fast_kernel
  double guard2;               // epoch kernels: guard band of the f64 count-based dE (second screening level), 1e-9 * Z * max|V|
  double fix_scale;            // word kernel: 2^-k, the unit of its fixed-point dE (word_metropolis.cuh)
  int m;
  int B[3], nb[3];
  int bxc, byc, bzc, box_sites;
  int n_modes;
  BrwBoxMode mode[BRW_MAX_MODES];
  int boxes_per_replica;
  int steps;
  int steps_a;                 // SPLIT word kernels: steps of warp group A (fewer warps, shorter steps: it gets more of them)
  int v_entries;               // S*S*n_shells
  int row_mul;                 // word kernel: row of its fixed-point table = code_a*row_mul + code_b
  // epoch kernels (epoch_metropolis.cuh): signed 8-bit digits of the fixed-point site-energy table, packed over the four
  // count fields, index (shell*3 + digit)*4 + species; guard band in fixed-point units
  int xdig[72];
  int gfix;
  int tma_stages;              // epoch kernel: 1 = box load through the bulk-async copy engine (TMA), 0 = LDG loop
  int h_rows;                  // cached rows per site: 4 (five species, relative to species 4) or 3 (relative to species 3)
};

struct BrwPlan {
  bool valid = false, use_box = false;
  int nbr_swap = 0;
  BrwBoxParams p{};
  int4 *d_classes = nullptr;   // residue classes o (x,y,z,unused)
  int4 *d_disp = nullptr;      // displacement classes d
  int *d_off = nullptr;        // [2][ztot] compact shared-memory offsets, by x-parity of the site
  double *d_Vrep = nullptr;    // [v_entries][16] lane-replicated V (layout [shell][centre][nbr])
  size_t smem = 0;
  int threads = 0, Mmax = 0;
  int M_a = 0;                 // SPLIT word kernels: trials per step of warp group A (runs p.steps_a steps)
  void *fast_fn = nullptr;     // specialised kernel for this (lattice, shells, pitch), if instantiated
  bool screened = false;
  bool split = false;          // word kernel with two warp groups per CTA and shared z margin planes
  bool pdl = false;            // fast_fn synchronises with the previous grid itself (griddepcontrol.wait): programmatic dependent launch
  bool byte_epoch = false;     // fast_fn is a byte-lattice epoch kernel (epoch_byte_metropolis.cuh)
  bool word = false;           // fast_fn is a word-lattice kernel (word_metropolis.cuh); d_Vrep holds its table blob
  size_t fast_smem = 0;
  // per-box counters
  unsigned long long *d_att = nullptr, *d_acc = nullptr;
  double *d_dE = nullptr;
  int n_slots = 0;
};

// ---- host planner -----------------------------------------------------------------------------
static inline int brw_posmod(int a, int m) { int r = a % m; return r < 0 ? r + m : r; }

// Period search.  For a period vector P = (Px,Py,Pz) (even entries) the residue classes are the
// lattice sites of the P box.  P is admissible if (a) no neighbour vector is = 0 (mod P) and (b) the set
// D of displacement classes d with no neighbour vector = +-d (mod P) is non-empty and connects all
// residue classes (so composition can flow between all sublattices).  Then the sites
// o + P*(i,j,k) and o + d + P*(i',j',k') are pairwise non-interacting.  In nbr_swap mode the second
// site is site1 + e (e a first-shell vector) and the requirement is that no neighbour vector equals
// P*k, P*k +- e for k != 0.
static bool brw_period_admissible(const BrwGeom &g, int nbr_swap, const std::vector<std::array<int, 3>> &first,
                                  const int P[3], std::vector<int4> &classes, std::vector<int4> &disp) {
  auto zero_mod = [&](int x, int y, int z) { return brw_posmod(x, P[0]) == 0 && brw_posmod(y, P[1]) == 0 && brw_posmod(z, P[2]) == 0; };
  for (int k = 0; k < g.ztot; k++)
    if (zero_mod(g.off[k][0], g.off[k][1], g.off[k][2])) return false;
  classes.clear();
  for (int z = 0; z < P[2]; z++) for (int y = 0; y < P[1]; y++) for (int x = 0; x < P[0]; x++)
    if (brw_is_site(g, x, y, z)) classes.push_back(make_int4(x, y, z, 0));
  disp.clear();
  if (nbr_swap) {
    // concurrent pairs (i, i+e) and (i', i'+e), i'-i = P*k != 0: differences P*k, P*k+-e
    for (auto &e : first) {
      if (zero_mod(e[0], e[1], e[2])) return false;
      for (int k = 0; k < g.ztot; k++)
        for (int sgn = -1; sgn <= 1; sgn++) {
          int vx = g.off[k][0] - sgn * e[0], vy = g.off[k][1] - sgn * e[1], vz = g.off[k][2] - sgn * e[2];
          if (vx == 0 && vy == 0 && vz == 0) continue;              // k = 0: the pair itself
          if (zero_mod(vx, vy, vz)) return false;
        }
      disp.push_back(make_int4(e[0], e[1], e[2], 0));
    }
    return true;
  }
  for (auto &c : classes) {
    if (c.x == 0 && c.y == 0 && c.z == 0) continue;
    bool good = true;
    for (int k = 0; k < g.ztot && good; k++)
      for (int sgn = -1; sgn <= 1; sgn += 2)
        if (zero_mod(g.off[k][0] - sgn * c.x, g.off[k][1] - sgn * c.y, g.off[k][2] - sgn * c.z)) { good = false; break; }
    if (good) disp.push_back(c);
  }
  if (disp.empty()) return false;
  // connectivity of the class graph under D
  std::vector<char> seen((size_t)P[0] * P[1] * P[2], 0);
  std::vector<int> stack{0};
  seen[0] = 1;
  size_t reached = 0;
  while (!stack.empty()) {
    int c = stack.back(); stack.pop_back(); reached++;
    int cx = c % P[0], cy = (c / P[0]) % P[1], cz = c / (P[0] * P[1]);
    for (auto &d : disp)
      for (int sgn = -1; sgn <= 1; sgn += 2) {
        int nx = brw_posmod(cx + sgn * d.x, P[0]), ny = brw_posmod(cy + sgn * d.y, P[1]), nz = brw_posmod(cz + sgn * d.z, P[2]);
        int n = (nz * P[1] + ny) * P[0] + nx;
        if (!seen[n]) { seen[n] = 1; stack.push_back(n); }
      }
  }
  return reached == classes.size();
}

struct BrwModeChoice { int P[3]; std::vector<int4> classes, disp; };
// Admissible periods in order of increasing volume: for each sorted triple every admissible orientation
// (up to BRW_MAX_MODES; phases cycle randomly through them so the move set stays isotropic).  Returns
// up to `max_sets` candidate sets; the caller scores them for the actual box.
static void brw_candidate_modes(const BrwGeom &g, int nbr_swap, const std::vector<std::array<int, 3>> &first, int Pmax,
                                bool cubic_only, int max_sets, std::vector<std::vector<BrwModeChoice>> &sets) {
  std::vector<std::array<int, 3>> triples;
  for (int a = 2; a <= Pmax; a += 2) for (int b = a; b <= Pmax; b += 2) for (int c = b; c <= Pmax; c += 2)
    if (!cubic_only || (a == b && b == c)) triples.push_back({a, b, c});
  std::sort(triples.begin(), triples.end(), [](const std::array<int, 3> &x, const std::array<int, 3> &y) {
    long vx = (long)x[0] * x[1] * x[2], vy = (long)y[0] * y[1] * y[2];
    return vx != vy ? vx < vy : x < y;
  });
  sets.clear();
  for (auto &t : triples) {
    std::array<int, 3> perm = t;
    std::vector<BrwModeChoice> modes;
    do {
      BrwModeChoice mc;
      mc.P[0] = perm[0]; mc.P[1] = perm[1]; mc.P[2] = perm[2];
      if ((int)modes.size() < BRW_MAX_MODES && brw_period_admissible(g, nbr_swap, first, mc.P, mc.classes, mc.disp))
        modes.push_back(std::move(mc));
    } while (std::next_permutation(perm.begin(), perm.end()));
    if (!modes.empty()) sets.push_back(std::move(modes));
    if ((int)sets.size() >= max_sets) break;
  }
}

// ---- kernel -------------------------------------------------------------------------------------
struct __align__(16) BrwStepParams {     // CTA-uniform per step, in shared memory
  int c1_base, c2_base;    // compact smem index of site1/site2 for (i,j,k) = 0
  int par1, par2;          // offset-table selector of site1 / site2
  int s[3];                // cyclic shift of the coarse index for site2
  int pad;
};

__device__ __forceinline__ int brw_box_compact(const BrwGeom &g, const BrwBoxParams &p, int X, int Y, int Z) {
  return (Z * p.byc + (Y >> g.ys)) * p.bxc + (X >> g.xs);
}
// offset-table selector: parity that decides how a doubled-grid dx maps to compact dx
__device__ __host__ __forceinline__ int brw_site_parity(const BrwGeom &g, int X, int Y, int Z) {
  (void)Y; (void)Z;
  return g.lattice == 0 ? 0 : (X & 1);   // bcc: x=y=z parity; fcc: x parity = (y+z) parity
}

template <int NBR>
__device__ __forceinline__ void brw_make_step(const BrwGeom &g, const BrwBoxParams &p, const BrwBoxMode &md,
                                              const int4 *classes, const int4 *disp, uint32_t k0, uint32_t k1,
                                              uint32_t step, uint32_t box_id, uint32_t phase_lo, BrwStepParams *out) {
  BrwPhilox4 r = brw_philox(0xFFFFFFFFu, step, box_id, phase_lo, k0, k1);
  int4 o = classes[md.cls0 + brw_below(r.x, md.n_classes)];
  int4 d = disp[md.d0 + brw_below(r.y, md.n_disp)];
  int X1 = p.m + o.x, Y1 = p.m + o.y, Z1 = p.m + o.z;
  int X2, Y2, Z2;
  if (NBR) { X2 = X1 + d.x; Y2 = Y1 + d.y; Z2 = Z1 + d.z; out->s[0] = out->s[1] = out->s[2] = 0; }
  else {
    X2 = p.m + (o.x + d.x) % md.P[0]; Y2 = p.m + (o.y + d.y) % md.P[1]; Z2 = p.m + (o.z + d.z) % md.P[2];
    out->s[0] = (int)brw_below(r.z, md.A[0]);
    out->s[1] = (int)(((r.w & 0xFFFFu) * (uint32_t)md.A[1]) >> 16);   // 16-bit draws; A <= 65535
    out->s[2] = (int)(((r.w >> 16) * (uint32_t)md.A[2]) >> 16);
  }
  out->c1_base = brw_box_compact(g, p, X1, Y1, Z1);
  out->c2_base = brw_box_compact(g, p, X2, Y2, Z2);
  out->par1 = brw_site_parity(g, X1, Y1, Z1);
  out->par2 = brw_site_parity(g, X2, Y2, Z2);
}

// sum of one site's shell chains for two centre species at once (the "before" centre and the
// "after" centre see the same neighbours): e_a, e_b in the reference's association.
template <int NBR>
__device__ __forceinline__ void brw_box_site_chains(const BrwGeom &g, const uint8_t *box, const int *off,
                                                    const double *Vl, int S, int c, int ca, int cb, int c_other,
                                                    int s_other_after, double &Ea, double &Eb) {
  // Vl points at lane slot (lane&15) of the replicated table [shell][centre][nbr][16]
  int k = 0;
  double ta = 0.0, tb = 0.0;
  for (int n = 0; n < g.n_shells; n++) {
    const double *Va = Vl + ((n * S + ca) * S) * 16;
    const double *Vb = Vl + ((n * S + cb) * S) * 16;
    double ea = 0.0, eb = 0.0;
    const int end = g.shell_end[n];
#pragma unroll 4
    for (; k < end; k++) {
      const int cn = c + off[k];
      int s = box[cn];
      ea = __dadd_rn(ea, Va[s * 16]);
      // in neighbour-swap mode the partner site is a neighbour and shows its new occupant
      if (NBR) s = (cn == c_other) ? s_other_after : s;
      eb = __dadd_rn(eb, Vb[s * 16]);
    }
    ta = (n == 0) ? ea : __dadd_rn(ta, ea);
    tb = (n == 0) ? eb : __dadd_rn(tb, eb);
  }
  Ea = ta; Eb = tb;
}

template <int NBR>
__global__ void __launch_bounds__(1024) brw_box_metropolis_kernel(
    BrwGeom g, BrwBoxParams p, uint8_t *__restrict__ lat, const double *__restrict__ beta,
    const double *__restrict__ Vrep, const int *__restrict__ off_g, const int4 *__restrict__ classes,
    const int4 *__restrict__ disp, uint32_t k0, uint32_t k1, uint32_t phase_lo, int mode,
    unsigned long long *__restrict__ att_out, unsigned long long *__restrict__ acc_out, double *__restrict__ dE_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const BrwBoxMode &md = p.mode[mode];
  double *Vs = reinterpret_cast<double *>(smem_raw);                       // [v_entries][16]
  int *off = reinterpret_cast<int *>(Vs + p.v_entries * 16);               // [2][ztot]
  BrwStepParams *sp = reinterpret_cast<BrwStepParams *>(off + 2 * g.ztot); // [2]
  double *red = reinterpret_cast<double *>(sp + 2);                        // [32]
  uint8_t *box = reinterpret_cast<uint8_t *>(red + 32);

  const int tid = threadIdx.x;
  const int replica = blockIdx.x / p.boxes_per_replica;
  const int bid = blockIdx.x - replica * p.boxes_per_replica;
  const int bi = bid % p.nb[0], bj = (bid / p.nb[0]) % p.nb[1], bk = bid / (p.nb[0] * p.nb[1]);
  uint8_t *L = lat + (long)replica * g.n_sites;

  // phase origin: even shift, identical for all boxes of a replica
  BrwPhilox4 ro = brw_philox(0xFFFFFFFEu, 0u, (uint32_t)replica, phase_lo, k0, k1);
  const int ox = 2 * (int)brw_below(ro.x, g.gx >> 1) + bi * p.B[0];
  const int oy = 2 * (int)brw_below(ro.y, g.gy >> 1) + bj * p.B[1];
  const int oz = 2 * (int)brw_below(ro.z, g.gz >> 1) + bk * p.B[2];

  for (int i = tid; i < p.v_entries * 16; i += blockDim.x) Vs[i] = Vrep[i];
  for (int i = tid; i < 2 * g.ztot; i += blockDim.x) off[i] = off_g[i];
  // ---- load box (global compact -> shared compact), rows along x are contiguous up to wrap
  for (int idx = tid; idx < p.box_sites; idx += blockDim.x) {
    int lxc = idx % p.bxc, t = idx / p.bxc, lyc = t % p.byc, lz = t / p.byc;
    int X, Y;
    if (g.lattice == 1) { X = 2 * lxc + (lz & 1); Y = 2 * lyc + (lz & 1); }
    else if (g.lattice == 2) { Y = lyc; X = 2 * lxc + ((lyc + lz) & 1); }
    else { X = lxc; Y = lyc; }
    int gxx = ox + X; if (gxx >= g.gx) gxx -= g.gx; if (gxx >= g.gx) gxx -= g.gx;
    int gyy = oy + Y; if (gyy >= g.gy) gyy -= g.gy; if (gyy >= g.gy) gyy -= g.gy;
    int gzz = oz + lz; if (gzz >= g.gz) gzz -= g.gz; if (gzz >= g.gz) gzz -= g.gz;
    box[idx] = L[brw_grid_to_compact(g, gxx, gyy, gzz)];
  }
  const uint32_t box_id = (uint32_t)blockIdx.x;
  if (tid == 0) brw_make_step<NBR>(g, p, md, classes, disp, k0, k1, 0u, box_id, phase_lo, &sp[0]);
  __syncthreads();

  const double my_beta = beta[replica];
  const double *Vl = Vs + (tid & 15);
  // strides of the coarse lattice in compact shared-memory index units
  const int stx = md.P[0] >> g.xs, sty = (md.P[1] >> g.ys) * p.bxc, stz = md.P[2] * p.byc * p.bxc;
  unsigned int n_att = 0, n_acc = 0;
  double dE_sum = 0.0;
  BrwPhilox4 rnd = {0, 0, 0, 0};

  for (int step = 0; step < p.steps; step++) {
    const BrwStepParams q = sp[step & 1];
    const int *off1 = off + q.par1 * g.ztot, *off2 = off + q.par2 * g.ztot;
    for (int t = tid; t < md.M; t += blockDim.x) {
      int i = t % md.A[0], r = t / md.A[0], j = r % md.A[1], k = r / md.A[1];
      int i2 = i + q.s[0]; if (i2 >= md.A[0]) i2 -= md.A[0];
      int j2 = j + q.s[1]; if (j2 >= md.A[1]) j2 -= md.A[1];
      int k2 = k + q.s[2]; if (k2 >= md.A[2]) k2 -= md.A[2];
      const int c1 = q.c1_base + i * stx + j * sty + k * stz;
      const int c2 = q.c2_base + i2 * stx + j2 * sty + k2 * stz;
      const int a = box[c1], b = box[c2];
      n_att++;
      if ((step & 3) == 0 || md.M > (int)blockDim.x)
        rnd = brw_philox((uint32_t)t, (uint32_t)step, box_id, phase_lo, k0, k1);
      if (a == b) { n_acc++; continue; }                       // src/metropolis.F90:774-777
      double E1a, E1b, E2b, E2a;
      brw_box_site_chains<NBR>(g, box, off1, Vl, g.S, c1, a, b, c2, a, E1a, E1b);
      brw_box_site_chains<NBR>(g, box, off2, Vl, g.S, c2, b, a, c1, b, E2b, E2a);
      const double before = __dadd_rn(E1a, E2b);               // pair_energy, sites unswapped
      const double after = __dadd_rn(E1b, E2a);                // pair_energy, sites swapped
      const double dE = __dsub_rn(after, before);              // :792
      bool accept = dE < 0.0;                                  // :796
      if (!accept) {
        const uint32_t w = (step & 3) == 0 ? rnd.x : (step & 3) == 1 ? rnd.y : (step & 3) == 2 ? rnd.z : rnd.w;
        accept = brw_u01(w) < exp(-my_beta * dE);              // :802
      }
      if (accept) { box[c1] = (uint8_t)b; box[c2] = (uint8_t)a; n_acc++; dE_sum += dE; }
    }
    if (tid == 0 && step + 1 < p.steps)
      brw_make_step<NBR>(g, p, md, classes, disp, k0, k1, (uint32_t)(step + 1), box_id, phase_lo, &sp[(step + 1) & 1]);
    __syncthreads();
  }

  // ---- store box
  for (int idx = tid; idx < p.box_sites; idx += blockDim.x) {
    int lxc = idx % p.bxc, t = idx / p.bxc, lyc = t % p.byc, lz = t / p.byc;
    int X, Y;
    if (g.lattice == 1) { X = 2 * lxc + (lz & 1); Y = 2 * lyc + (lz & 1); }
    else if (g.lattice == 2) { Y = lyc; X = 2 * lxc + ((lyc + lz) & 1); }
    else { X = lxc; Y = lyc; }
    int gxx = ox + X; if (gxx >= g.gx) gxx -= g.gx; if (gxx >= g.gx) gxx -= g.gx;
    int gyy = oy + Y; if (gyy >= g.gy) gyy -= g.gy; if (gyy >= g.gy) gyy -= g.gy;
    int gzz = oz + lz; if (gzz >= g.gz) gzz -= g.gz; if (gzz >= g.gz) gzz -= g.gz;
    L[brw_grid_to_compact(g, gxx, gyy, gzz)] = box[idx];
  }
  // ---- counters: fixed-order CTA reduction, one slot per box (deterministic)
  unsigned int packed_att = n_att, packed_acc = n_acc;
  for (int o = 16; o > 0; o >>= 1) {
    packed_att += __shfl_down_sync(0xffffffffu, packed_att, o);
    packed_acc += __shfl_down_sync(0xffffffffu, packed_acc, o);
    dE_sum += __shfl_down_sync(0xffffffffu, dE_sum, o);
  }
  __shared__ unsigned int s_att[32], s_acc[32];
  if ((tid & 31) == 0) { s_att[tid >> 5] = packed_att; s_acc[tid >> 5] = packed_acc; red[tid >> 5] = dE_sum; }
  __syncthreads();
  if (tid == 0) {
    unsigned long long A = 0, C = 0; double D = 0.0;
    for (int w = 0; w < (int)((blockDim.x + 31) >> 5); w++) { A += s_att[w]; C += s_acc[w]; D += red[w]; }
    att_out[blockIdx.x] += A; acc_out[blockIdx.x] += C; dE_out[blockIdx.x] += D;
  }
}

// ---- specialised kernel: compile-time geometry ---------------------------------------------------
// Same algorithm as brw_box_metropolis_kernel<0>, instantiated for a fixed (lattice, n_shells, box
// pitch): every neighbour offset is an immediate of the LDS that gathers it, the shell loops are
// fully unrolled in the reference's summation order, and the lane-replicated V table is addressed
// with one shift-add per lookup.  ~2x fewer issued instructions than the generic kernel.
#include <utility>
template <int LAT> struct BrwTab;
template <> struct BrwTab<1> {
  static constexpr int start(int n) { return brw_bcc_start[n]; }
  static constexpr int count(int n) { return brw_bcc_count[n]; }
  static constexpr int off(int k, int c) { return brw_bcc_off[k][c]; }
};
template <> struct BrwTab<2> {
  static constexpr int start(int n) { return brw_fcc_start[n]; }
  static constexpr int count(int n) { return brw_fcc_count[n]; }
  static constexpr int off(int k, int c) { return brw_fcc_off[k][c]; }
};
template <int LAT, int N> struct BrwShellRange {
  static constexpr int start = BrwTab<LAT>::start(N), count = BrwTab<LAT>::count(N);
};
constexpr int brw_fdiv2(int v) { return v >= 0 ? v / 2 : -((-v + 1) / 2); }
// compact shared-memory offset of neighbour K for a centre site of x-parity PAR
template <int LAT, int PX, int PY, int PAR, int K>
struct BrwCOff {
  static constexpr int dx = BrwTab<LAT>::off(K, 0), dy = BrwTab<LAT>::off(K, 1), dz = BrwTab<LAT>::off(K, 2);
  static constexpr int dxc = brw_fdiv2(PAR + dx);
  static constexpr int dyc = LAT == 1 ? brw_fdiv2(PAR + dy) : dy;
  static constexpr int value = (dz * PY + dyc) * PX + dxc;
};
template <int LAT, int PX, int PY, int PAR, int K0, int... Is>
__device__ __forceinline__ void brw_fast_shell(const uint8_t *bc, const char *Va, const char *Vb, double &ea,
                                               double &eb, std::integer_sequence<int, Is...>) {
  // comma fold keeps the reference's left-to-right neighbour order
  // box bytes hold 8*species, so (byte << 4) is the 128-byte stride of the lane-replicated V table
  ((ea = __dadd_rn(ea, *reinterpret_cast<const double *>(Va + ((int)bc[BrwCOff<LAT, PX, PY, PAR, K0 + Is>::value] << 4))),
    eb = __dadd_rn(eb, *reinterpret_cast<const double *>(Vb + ((int)bc[BrwCOff<LAT, PX, PY, PAR, K0 + Is>::value] << 4)))),
   ...);
}
// ---- integer neighbour counts (screening path) ----------------------------------------------------
// acc gets one 8-bit field per species 0..3: field s counts neighbours of species s in this shell.
// PTX shl.b32 clamps shift amounts > 31 to "all bits out", so species 4 (byte 32) adds 0 and is
// recovered as Z_n - sum(others).  Counts are <= 32 per shell, so fields never overflow.
__device__ __forceinline__ uint32_t brw_shl1(uint32_t sh) {
  uint32_t r;
  asm("shl.b32 %0, 1, %1;" : "=r"(r) : "r"(sh));
  return r;
}
// ld.shared.u8 straight into a 32-bit register with the neighbour offset as an immediate (avoids the
// byte->word PRMT the C++ uint8_t load would add).  Neighbours are never the two sites of the
// trial, so ordering against this thread's own box stores is irrelevant.
template <int OFF>
__device__ __forceinline__ uint32_t brw_lds_u8(uint32_t saddr) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1+%2];" : "=r"(v) : "r"(saddr), "n"(OFF));
  return v;
}
template <int LAT, int PX, int PY, int PAR, int K0, int... Is>
__device__ __forceinline__ uint32_t brw_count_shell(uint32_t saddr, std::integer_sequence<int, Is...>) {
  return (brw_shl1(brw_lds_u8<BrwCOff<LAT, PX, PY, PAR, K0 + Is>::value>(saddr)) + ...);
}
template <int LAT, int NSH, int PX, int PY, int PAR, int N>
__device__ __forceinline__ void brw_count_shells(uint32_t saddr, uint32_t (&acc)[NSH]) {
  if constexpr (N < NSH) {
    acc[N] = brw_count_shell<LAT, PX, PY, PAR, BrwShellRange<LAT, N>::start>(
        saddr, std::make_integer_sequence<int, BrwShellRange<LAT, N>::count>{});
    brw_count_shells<LAT, NSH, PX, PY, PAR, N + 1>(saddr, acc);
  }
}
template <int LAT, int NSH, int PX, int PY, int PAR, int N>
__device__ __forceinline__ void brw_fast_shells(const uint8_t *bc, const char *Vl, int S, int ca, int cb, double &Ea,
                                                double &Eb) {
  if constexpr (N < NSH) {
    const char *Va = Vl + ((N * S + ca) * S) * 128, *Vb = Vl + ((N * S + cb) * S) * 128;
    double ea = 0.0, eb = 0.0;
    brw_fast_shell<LAT, PX, PY, PAR, BrwShellRange<LAT, N>::start>(
        bc, Va, Vb, ea, eb, std::make_integer_sequence<int, BrwShellRange<LAT, N>::count>{});
    if constexpr (N == 0) { Ea = ea; Eb = eb; }
    else { Ea = __dadd_rn(Ea, ea); Eb = __dadd_rn(Eb, eb); }
    brw_fast_shells<LAT, NSH, PX, PY, PAR, N + 1>(bc, Vl, S, ca, cb, Ea, Eb);
  }
}

// ---- tiled full-lattice energy (production total_energy, src/bw_hamiltonian.f90:58-81) ----------------
// E = 1/2 sum_sites nbr_energy(site).  One CTA per 32 x 16 x 8 tile of the compact lattice: the tile and its halo
// (periodic wrap resolved while staging) go to shared memory as 8*species bytes, then every thread forms the
// integer neighbour counts of its sites with the compile-time gathers of the screened Metropolis kernels
// (one LDS.U8 with an immediate offset + shl + add per neighbour) and
//     nbr_energy = sum_shell sum_b count[shell][b] * V_shell(centre, b)
// in f64 (species 4 inferred from the coordination number).  Deterministic (fixed site -> thread -> tile order);
// equal to the reference's sequential sum up to f64 rounding -- the bit-exact order is brw_ordered_sum_kernel.
// Algorithmic bytes: N read (1 B/site); staged bytes: 2.8 B/site (halo); ~6x fewer instructions per site than
// the generic row kernel (brw_energy_partial_kernel), which remains the path for every other geometry.
template <int LAT> struct BrwETile {
  static constexpr int TX = 32, TY = 16, TZ = 8, HX = 2, HY = LAT == 1 ? 2 : 4, HZ = 4;
  static constexpr int PX = TX + 2 * HX, PY = TY + 2 * HY, PZ = TZ + 2 * HZ;
};
template <int LAT, int NSH>
__global__ void __launch_bounds__(256) brw_energy_tile_kernel(BrwGeom g, const double *__restrict__ V,
                                                              const uint8_t *__restrict__ lat,
                                                              double *__restrict__ partial, int ntx, int nty) {
  using T = BrwETile<LAT>;
  __shared__ __align__(16) uint8_t box[T::PZ * T::PY * T::PX];
  __shared__ double Vs[BRW_MAX_SHELLS * 25];
  __shared__ double red[8];
  const int tid = threadIdx.x;
  const int tile = blockIdx.x, tx = tile % ntx, ty = (tile / ntx) % nty, tz = tile / (ntx * nty);
  const uint8_t *L = lat + (long)blockIdx.y * g.n_sites;
  const int x0 = tx * T::TX - T::HX, y0 = ty * T::TY - T::HY, z0 = tz * T::TZ - T::HZ;
  for (int i = tid; i < g.S * g.S * NSH; i += 256) Vs[i] = V[i];
  for (int i = tid; i < T::PZ * T::PY * T::PX; i += 256) {
    const int lx = i % T::PX, ly = (i / T::PX) % T::PY, lz = i / (T::PX * T::PY);
    int xc = x0 + lx, yc = y0 + ly, z = z0 + lz;
    xc += xc < 0 ? g.cx : 0; xc -= xc >= g.cx ? g.cx : 0;
    yc += yc < 0 ? g.cy : 0; yc -= yc >= g.cy ? g.cy : 0;
    z += z < 0 ? g.cz : 0; z -= z >= g.cz ? g.cz : 0;
    box[i] = (uint8_t)(L[((long)z * g.cy + yc) * g.cx + xc] << 3);
  }
  __syncthreads();
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(box);
  const int S = g.S;
  double acc = 0.0;
#pragma unroll 1
  for (int i = tid; i < T::TX * T::TY * T::TZ; i += 256) {
    const int lx = i % T::TX, ly = (i / T::TX) % T::TY, lz = i / (T::TX * T::TY);      // a warp = one x-row
    const int off = ((lz + T::HZ) * T::PY + ly + T::HY) * T::PX + lx + T::HX;
    const int par = LAT == 1 ? (lz & 1) : ((ly + lz) & 1);                             // tile origins are even
    uint32_t cnt[NSH];
    if (par) brw_count_shells<LAT, NSH, T::PX, T::PY, 1, 0>(sbase + off, cnt);
    else brw_count_shells<LAT, NSH, T::PX, T::PY, 0, 0>(sbase + off, cnt);
    const double *Va = Vs + (box[off] >> 3);                                           // V(centre, nbr, shell)
    double e = 0.0;
#pragma unroll
    for (int n = 0; n < NSH; n++) {
      const uint32_t c = cnt[n];
      int rest = g.shell_end[n] - (n ? g.shell_end[n - 1] : 0);        // coordination number of the shell
#pragma unroll
      for (int b = 0; b < 4; b++) {
        if (b < S) {
          const int cb = (int)((c >> (8 * b)) & 255u);
          rest -= cb;
          e += (double)cb * Va[(n * S + b) * S];
        }
      }
      if (S == 5) e += (double)rest * Va[(n * S + 4) * S];
    }
    acc += e;
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if ((tid & 31) == 0) red[tid >> 5] = acc;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; w++) t += red[w];
    partial[(long)blockIdx.y * gridDim.x + tile] = t;
  }
}

// box <-> global copy: one warp per compact-x row (PX consecutive bytes, wrapping inside the
// global row), rows unrolled x8 so the independent loads overlap.  STORE=false: global -> shared.
// PITCH: bytes between consecutive x-rows of the shared-memory box (>= PX; padded where the gathers would otherwise hit
// the same banks from the 2-3 rows a warp covers).
template <int LAT, int PX, int PY, bool STORE, int PITCH = PX>
__device__ __forceinline__ void brw_box_copy(const BrwGeom &g, uint8_t *L, uint8_t *box, int n_rows, int ox, int oy,
                                             int oz) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  static_assert(PX <= 32, "one warp covers a row");
  const int oxc = ox >> 1;                      // origin is even: compact x origin
#pragma unroll 8
  for (int r = warp; r < n_rows; r += nwarps) {
    const int lyc = r % PY, lz = r / PY;
    int gzz = oz + lz; if (gzz >= g.gz) gzz -= g.gz; if (gzz >= g.gz) gzz -= g.gz;
    const int Y = LAT == 1 ? 2 * lyc + (lz & 1) : lyc;
    int gyy = oy + Y; if (gyy >= g.gy) gyy -= g.gy; if (gyy >= g.gy) gyy -= g.gy;
    if (lane < PX) {
      int gxc = oxc + lane; if (gxc >= g.cx) gxc -= g.cx; if (gxc >= g.cx) gxc -= g.cx;
      const long gi = ((long)gzz * g.cy + (gyy >> g.ys)) * g.cx + gxc;
      if (STORE) L[gi] = (uint8_t)(box[r * PITCH + lane] >> 3);  // shared-memory bytes hold 8*species
      else box[r * PITCH + lane] = (uint8_t)(L[gi] << 3);
    }
  }
}

// SCREEN = true: the decision is first attempted with dE formed from integer neighbour counts
// (dE = sum_n sum_s (c1-c2)[n][s] * (V_n[b][s] - V_n[a][s]); exact counts, different f64 rounding).
// Whenever that value is within a guard band of a decision boundary (|dE| < guard, or |u - exp(-beta dE)|
// within the propagated band) the trial is recomputed with the reference's association and decided
// by it.  The guard (host: 1e-9 * ztot * max|V|) exceeds the worst-case rounding difference between
// the two formulas by > 4 orders of magnitude, so every accept/reject equals the one the reference
// arithmetic yields: trajectories are identical to SCREEN = false
// (test_screened_kernel_trajectory_identical).
template <int LAT, int NSH, int PX, int PY, bool SCREEN, int MAXT>
__global__ void __launch_bounds__(MAXT) brw_box_metropolis_fast_kernel(
    BrwGeom g, BrwBoxParams p, uint8_t *__restrict__ lat, const double *__restrict__ beta,
    const double *__restrict__ Vrep, const int4 *__restrict__ classes, const int4 *__restrict__ disp, uint32_t k0,
    uint32_t k1, uint32_t phase_lo, int mode, unsigned long long *__restrict__ att_out,
    unsigned long long *__restrict__ acc_out, double *__restrict__ dE_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const BrwBoxMode &md = p.mode[mode];
  double *Vs = reinterpret_cast<double *>(smem_raw);                       // [v_entries][16]
  double *red = Vs + p.v_entries * 16;                                     // [32]
  BrwStepParams *sp = reinterpret_cast<BrwStepParams *>(red + 32);         // [steps]
  uint8_t *box = reinterpret_cast<uint8_t *>(sp + p.steps);                // [bzc][PY][PX]
  __shared__ unsigned int s_att[32], s_acc[32];

  const int tid = threadIdx.x;
  const int replica = blockIdx.x / p.boxes_per_replica;
  const int bid = blockIdx.x - replica * p.boxes_per_replica;
  const int bi = bid % p.nb[0], bj = (bid / p.nb[0]) % p.nb[1], bk = bid / (p.nb[0] * p.nb[1]);
  uint8_t *L = lat + (long)replica * g.n_sites;

  BrwPhilox4 ro = brw_philox(0xFFFFFFFEu, 0u, (uint32_t)replica, phase_lo, k0, k1);
  const int ox = 2 * (int)brw_below(ro.x, g.gx >> 1) + bi * p.B[0];
  const int oy = 2 * (int)brw_below(ro.y, g.gy >> 1) + bj * p.B[1];
  const int oz = 2 * (int)brw_below(ro.z, g.gz >> 1) + bk * p.B[2];
  const uint32_t box_id = (uint32_t)blockIdx.x;

  for (int i = tid; i < p.v_entries * 16; i += blockDim.x) Vs[i] = Vrep[i];
  // every step's CTA-uniform parameters, computed in parallel up front (one thread per step)
  for (int st = tid; st < p.steps; st += blockDim.x)
    brw_make_step<0>(g, p, md, classes, disp, k0, k1, (uint32_t)st, box_id, phase_lo, &sp[st]);
  brw_box_copy<LAT, PX, PY, false>(g, L, box, PY * p.bzc, ox, oy, oz);
  __syncthreads();

  const double my_beta = beta[replica];
  const char *Vl = reinterpret_cast<const char *>(Vs + (tid & 15));
  const uint32_t box_s = (uint32_t)__cvta_generic_to_shared(box);
  const int S = g.S;
  const int stx = md.P[0] >> 1, sty = (LAT == 1 ? (md.P[1] >> 1) : md.P[1]) * PX, stz = md.P[2] * PY * PX;
  // this thread's coarse cell (fixed for the whole phase; blockDim >= M is guaranteed by the host)
  const bool active = tid < md.M;
  const int A0 = md.A[0], A1 = md.A[1], A2 = md.A[2];
  const int ci = tid % A0, cr = tid / A0, cj = cr % A1, ck = cr / A1;
  const int base1 = ci * stx + cj * sty + ck * stz;
  unsigned int n_att = 0, n_acc = 0;
  double dE_sum = 0.0;
  BrwPhilox4 rnd = {0, 0, 0, 0};

  for (int step = 0; step < p.steps; step++) {
    const BrwStepParams q = sp[step];
    if (active) {
      int i2 = ci + q.s[0]; if (i2 >= A0) i2 -= A0;
      int j2 = cj + q.s[1]; if (j2 >= A1) j2 -= A1;
      int k2 = ck + q.s[2]; if (k2 >= A2) k2 -= A2;
      const int c1 = q.c1_base + base1;
      const int c2 = q.c2_base + i2 * stx + j2 * sty + k2 * stz;
      const int a = box[c1], b = box[c2];
      n_att++;
      if ((step & 3) == 0) rnd = brw_philox((uint32_t)tid, (uint32_t)step, box_id, phase_lo, k0, k1);
      if (a != b) {
        const int sa = a >> 3, sb = b >> 3;                    // species (box bytes hold 8*species)
        const uint32_t w = (step & 3) == 0 ? rnd.x : (step & 3) == 1 ? rnd.y : (step & 3) == 2 ? rnd.z : rnd.w;
        const double u = brw_u01(w);
        bool decided = false, accept = false;
        double dE = 0.0;
        if (SCREEN) {
          uint32_t A1[NSH], A2[NSH];
          const uint32_t sa1 = box_s + c1, sa2 = box_s + c2;
          if (q.par1) brw_count_shells<LAT, NSH, PX, PY, 1, 0>(sa1, A1);
          else brw_count_shells<LAT, NSH, PX, PY, 0, 0>(sa1, A1);
          if (q.par2) brw_count_shells<LAT, NSH, PX, PY, 1, 0>(sa2, A2);
          else brw_count_shells<LAT, NSH, PX, PY, 0, 0>(sa2, A2);
          // fully unrolled over (shell, species): counts stay in registers
#pragma unroll
          for (int n = 0; n < NSH; n++) {
            const double *Va = reinterpret_cast<const double *>(Vl + ((n * S + sa) * S) * 128);
            const double *Vb = reinterpret_cast<const double *>(Vl + ((n * S + sb) * S) * 128);
            const uint32_t x1 = A1[n], x2 = A2[n];
            int rest = 0;
#pragma unroll
            for (int sp = 0; sp < 4; sp++) {
              if (sp < S) {
                const int d = (int)__byte_perm(x1, 0, 0x4440 | sp) - (int)__byte_perm(x2, 0, 0x4440 | sp);
                rest -= d;
                dE = fma((double)d, Vb[sp * 16] - Va[sp * 16], dE);
              }
            }
            if (S == 5) dE = fma((double)rest, Vb[4 * 16] - Va[4 * 16], dE);
          }
          if (fabs(dE) > p.guard) {
            if (dE < 0.0) { accept = true; decided = true; }
            else {
              const double t = exp(-my_beta * dE);
              if (fabs(u - t) > t * (my_beta * p.guard + 1e-12)) { accept = u < t; decided = true; }
            }
          }
        }
        if (!decided) {
          double E1a, E1b, E2b, E2a;
          if (q.par1) brw_fast_shells<LAT, NSH, PX, PY, 1, 0>(box + c1, Vl, S, sa, sb, E1a, E1b);
          else brw_fast_shells<LAT, NSH, PX, PY, 0, 0>(box + c1, Vl, S, sa, sb, E1a, E1b);
          if (q.par2) brw_fast_shells<LAT, NSH, PX, PY, 1, 0>(box + c2, Vl, S, sb, sa, E2b, E2a);
          else brw_fast_shells<LAT, NSH, PX, PY, 0, 0>(box + c2, Vl, S, sb, sa, E2b, E2a);
          const double before = __dadd_rn(E1a, E2b);           // pair_energy, sites unswapped
          const double after = __dadd_rn(E1b, E2a);            // pair_energy, sites swapped
          dE = __dsub_rn(after, before);                       // src/metropolis.F90:792
          accept = dE < 0.0;                                   // :796
          if (!accept) accept = u < exp(-my_beta * dE);        // :802
        }
        if (accept) { box[c1] = (uint8_t)b; box[c2] = (uint8_t)a; n_acc++; dE_sum += dE; }
      } else n_acc++;                                          // :774-777
    }
    __syncthreads();
  }

  brw_box_copy<LAT, PX, PY, true>(g, L, box, PY * p.bzc, ox, oy, oz);
  for (int o = 16; o > 0; o >>= 1) {
    n_att += __shfl_down_sync(0xffffffffu, n_att, o);
    n_acc += __shfl_down_sync(0xffffffffu, n_acc, o);
    dE_sum += __shfl_down_sync(0xffffffffu, dE_sum, o);
  }
  if ((tid & 31) == 0) { s_att[tid >> 5] = n_att; s_acc[tid >> 5] = n_acc; red[tid >> 5] = dE_sum; }
  __syncthreads();
  if (tid == 0) {
    unsigned long long A = 0, C = 0; double D = 0.0;
    for (int w = 0; w < (int)((blockDim.x + 31) >> 5); w++) { A += s_att[w]; C += s_acc[w]; D += red[w]; }
    att_out[blockIdx.x] += A; acc_out[blockIdx.x] += C; dE_out[blockIdx.x] += D;
  }
}
