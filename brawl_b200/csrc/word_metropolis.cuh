// word_metropolis.cuh -- production Metropolis (nbr_energy,
// pair_energy, pair_swap, monte_carlo_step_lattice; src/metropolis.F90:751-813) on a WORD lattice
// with DENSE non-interacting site sets.
//
// Lattice in shared memory: one 32-bit word per site, word = 1 << (8*species) for species 0..3 and 0
// for species 4.  The integer neighbour counts of the screened dE then need ONE LDS.32 + half an
// IADD3 per neighbour:
//     acc[shell] = 0x80808080 + sum_{nbrs of site1} word - sum_{nbrs of site2} word
// leaves byte s of acc[shell] = 128 + (c1 - c2)[shell][s] (|c1-c2| <= 24, so no borrow crosses a
// byte), and
//     dE = sum_shell sum_s (c1-c2)[shell][s] * U[a][b][shell][s],   U = (V[b][s]-V[a][s]) - (V[b][4]-V[a][4])
// is evaluated in FIXED POINT with dp4a: U*2^k is rounded to an integer and split into NLIMB signed
// 8-bit digits, so one dp4a.u32.s32 per (shell, digit) multiplies the four byte fields at once.  The
// rounding error is bounded rigorously by Z*2^-k (sum_f |d_f| <= 2Z, half a unit each) and is part
// of the guard band; any trial whose dE or acceptance test lies inside the band is recomputed with
// the reference's f64 association (src/bw_hamiltonian.f90:171-173, :1014-1017, :111-112;
// src/metropolis.F90:792-802) and decided by it, exactly like the byte-lattice screened kernel.  The
// acceptance test uses ex2.approx.f32 first (relative error < 1e-5 over the non-flushed range) with
// a correspondingly wider band.  EXACT = true skips the screening: every trial is decided by the
// reference association (the trajectory-identity tests compare the two instantiations).
//
// Decomposition (per step, CTA-uniform): the densest set of mutually NON-INTERACTING sites of the
// bcc lattice with <= 4 shells is  S_o = o + {(0,0,0),(2,2,2)} + 4 Z^3  (1/8 of all sites: no
// neighbour vector of shells 1-4 is = (0,0,0) or (2,2,2) mod 4 -- checked by the host planner).  A
// step draws two residue classes o, o' (mod 4) and proposes, in parallel, swaps between sites of S_o in
// the lower-y half of the box's active region and sites of S_o' in the upper-y half; the two halves
// are separated by a gap >= the interaction reach, so all 2M sites of a step are pairwise
// non-interacting for ANY o, o' and the M simultaneous Metropolis decisions are exactly equivalent to
// M sequential ones.  Every pair of residue classes is connected (composition flows between all
// sublattices), each move is its own inverse and its proposal probability does not depend on the
// configuration: detailed balance holds move by move.
//
// Lane mapping: a warp takes one x-row of sub-class A sites (lanes 0..13, even word addresses) and
// one x-row of sub-class B = A + (2,2,2) sites (lanes 14..27, odd word addresses): consecutive sites of
// a row are 2 words apart, so the 28 lanes fall into 28 different banks for every gather.  Site 2 is
// the same construction in the upper half with a random row rotation, a cyclic x shift and an
// optional A<->B exchange (all CTA-uniform), which keeps its gathers conflict-free as well.
//
// SPLIT: the rows of a box are divided between two warp groups on disjoint z zones (a frozen gap plane of 4 grid
// units >= the interaction reach between them), each with its own named barrier, so a step of one group (gather,
// arithmetic, barrier wait, the rare reference-association trial) does not stall the other.  Boxes of such a
// kernel are 32 planes deep at a layer pitch of 28: consecutive layers share their frozen margin planes.
#pragma once
#include "tile_metropolis.cuh"

template <int LAT, int PXP, int PLP, int PAR, int K>
struct BrwWOff {
  static constexpr int dx = BrwTab<LAT>::off(K, 0), dy = BrwTab<LAT>::off(K, 1), dz = BrwTab<LAT>::off(K, 2);
  static constexpr int dxc = brw_fdiv2(PAR + dx);
  static constexpr int dyc = LAT == 1 ? brw_fdiv2(PAR + dy) : dy;
  static constexpr int value = dz * PLP + dyc * PXP + dxc;      // in words
};
template <int LAT, int PXP, int PLP, int PAR, int K0, int... Is>
__device__ __forceinline__ uint32_t brw_wsum_shell(const uint32_t *wc, std::integer_sequence<int, Is...>) {
  return (wc[BrwWOff<LAT, PXP, PLP, PAR, K0 + Is>::value] + ...);
}
// acc[n] += sum over shell n of the neighbour words of the site at wc
template <int LAT, int NSH, int PXP, int PLP, int PAR, int N>
__device__ __forceinline__ void brw_wsum_shells(const uint32_t *wc, uint32_t (&acc)[NSH]) {
  if constexpr (N < NSH) {
    acc[N] = brw_wsum_shell<LAT, PXP, PLP, PAR, BrwShellRange<LAT, N>::start>(
        wc, std::make_integer_sequence<int, BrwShellRange<LAT, N>::count>{});
    brw_wsum_shells<LAT, NSH, PXP, PLP, PAR, N + 1>(wc, acc);
  }
}
__device__ __forceinline__ int brw_dp4a_us(uint32_t a, int b, int c) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ float brw_ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// species code of a site word: clz>>3 = 3,2,1,0 for species 0..3 and 4 for species 4 (word 0)
__device__ __forceinline__ int brw_word_code(uint32_t w) { return __clz((int)w) >> 3; }
__device__ __forceinline__ int brw_code_species(int code) { return code == 4 ? 4 : 3 - code; }
__host__ __device__ __forceinline__ uint32_t brw_species_word(int s) { return s < 4 ? 1u << (8 * s) : 0u; }

// ---- pair words (PAIRW): W[c] = nibbles(site c) | nibbles(site c+1) << 16, nibbles = 1 << 4*species (species 4: 0).
// One LDS.32 returns two x-adjacent sites for ANY c, so 30 loads cover the 50 neighbours (pair_gather.inc, generated
// by tools/gen_pair_gather.py; its plan is re-derived and emulated on the CPU in tests/test_dense_plan.py).
__host__ __device__ __forceinline__ uint32_t brw_species_nibbles(int s) { return s < 4 ? 1u << (4 * s) : 0u; }
// species code of the low lane: 3,2,1,0 for species 0..3 and 4 for species 4 (same codes as brw_word_code)
__device__ __forceinline__ int brw_pair_code(uint32_t w) { return (__clz((int)(w & 0xFFFFu)) - 16) >> 2; }
// the same code without the XU pipe (FLO): (w * 0x1234) >> 12 & 7 = species + 1 for the one-hot nibbles 1, 16, 256, 4096 and
// 0 for species 4 (w = 0); 4 - that = brw_pair_code(w).  One IMAD + SHF + LOP3 + IADD on the FMA/ALU pipes.
__device__ __forceinline__ int brw_pair_code_alu(uint32_t w16) { return 4 - (int)(((w16 * 0x1234u) >> 12) & 7u); }
// four 4-bit fields of the low 16 bits -> four 8-bit fields
__device__ __forceinline__ uint32_t brw_nib2byte(uint32_t x) {
  const uint32_t t = __byte_perm(x, 0u, 0x4140);          // [b0, 0, b1, 0]
  return (t | (t << 4)) & 0x0F0F0F0Fu;
}
__device__ __forceinline__ uint32_t brw_lo16(const uint32_t *p) { return *reinterpret_cast<const uint16_t *>(p); }
#include "pair_gather.inc"
#ifndef BRW_XP
#define BRW_XP 0        // timing experiments only (results invalid): 1 skips the dE arithmetic, 2 skips the gathers
#endif
#ifndef BRW_ANTI
#define BRW_ANTI 0      // 1: SPLIT groups in enforced anti-phase (timing experiment)
#endif
#ifndef BRW_INTDEC
#define BRW_INTDEC 0    // 1: acceptance test without conversions / MUFU / fp64 (FMA + ALU pipes only); measured slower
#endif
#ifndef BRW_TABG
#define BRW_TABG 0      // 1: the fixed-point table row is loaded in the gather half (costs registers across the barrier)
#endif

// box <-> global copy, one warp per compact-x row of PX sites.  STORE=false: global bytes -> shared words.
template <int LAT, int PX, int PY, int PXP, int PLP, bool STORE, bool PAIRW>
__device__ __forceinline__ void brw_wbox_copy(const BrwGeom &g, uint8_t *L, uint32_t *wbox, int row_begin, int n_rows,
                                              int ox, int oy, int oz) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  static_assert(PX <= 32 && (!PAIRW || PX == 32), "one warp covers a row");
  const int oxc = ox >> 1;
#pragma unroll 8
  for (int r = row_begin + warp; r < n_rows; r += nwarps) {
    const int lyc = r % PY, lz = r / PY;
    int gzz = oz + lz; if (gzz >= g.gz) gzz -= g.gz; if (gzz >= g.gz) gzz -= g.gz;
    const int Y = LAT == 1 ? 2 * lyc + (lz & 1) : lyc;
    int gyy = oy + Y; if (gyy >= g.gy) gyy -= g.gy; if (gyy >= g.gy) gyy -= g.gy;
    if (lane < PX) {
      int gxc = oxc + lane; if (gxc >= g.cx) gxc -= g.cx; if (gxc >= g.cx) gxc -= g.cx;
      const long gi = ((long)gzz * g.cy + (gyy >> g.ys)) * g.cx + gxc;
      uint32_t *w = wbox + lz * PLP + lyc * PXP + lane;
      if (PAIRW) {
        if (STORE) L[gi] = (uint8_t)brw_code_species(brw_pair_code(*w));
        else {
          const uint32_t v = brw_species_nibbles(L[gi]);
          uint32_t hi = __shfl_down_sync(0xffffffffu, v, 1);    // the x-neighbour's nibbles (PX == 32: all lanes active)
          if (lane == 31) hi = 0u;                              // beyond the row: never a useful lane (margin)
          *w = v | (hi << 16);
        }
      } else {
        if (STORE) L[gi] = (uint8_t)brw_code_species(brw_word_code(*w));
        else *w = brw_species_word(L[gi]);
      }
    }
  }
}

// Constant tables of the word kernel, one blob in global memory copied to shared memory per CTA:
//   int32  urow[PAIRS+1][32][2]  fixed-point table, pair-major so that different rows (species pairs) fall into
//                           different banks: element e = n*NLIMB + k of row r (r = code_a*row_mul + code_b) is word
//                           [e/2][r][e%2] = the k-th signed digit of U*2^k for the four species fields packed as
//                           int8x4; pair PAIRS = int64 K = 128 * sum_f Ufix_f
//   int32  off[2][ztot]     word offsets of the neighbours per x-parity (reference-association path)
//   double V[n_shells][S][S] reference V_ex in Fortran order V(centre, nbr, shell)
template <int NSH, int NLIMB> struct BrwWordTab {
  static constexpr int PAIRS = NSH * NLIMB / 2;
  static constexpr int urow_words = (PAIRS + 1) * 64;
};

// geometry of the dense decomposition for one box (compile-time)
template <int BX, int BY, int BZ, int MARGIN> struct BrwDenseGeom {
  static constexpr int NI = (BX - 2 * MARGIN) / 4;            // sites per x-row of one sub-class (14)
  static constexpr int HALF = ((BY - 2 * MARGIN - MARGIN) / 2) & ~3;   // y extent of one half, gap >= MARGIN (24)
  static constexpr int NJ = HALF / 4;                          // rows per plane and half (6)
  static constexpr int NP = (BZ - 2 * MARGIN) / 4;             // planes per sub-class (6)
  static constexpr int NROWS = NJ * NP;                        // rows per sub-class and half (36)
  static constexpr int Y_UPPER = MARGIN + HALF + MARGIN;       // first y of the upper half (32)
  static_assert(2 * NI <= 32, "one warp holds an A row and a B row");
  // SPLIT: two independent warp groups on disjoint z zones (planes [0, NPA) and [NPA+1, NP); plane NPA is a frozen
  // gap of 4 grid units >= the interaction reach), each with its own named barrier
  static constexpr int NPA = (NP - 1) / 2, NPB = NP - 1 - NPA;
  static constexpr int NRA = NJ * NPA, NRB = NJ * NPB;
};

struct __align__(16) BrwDenseStep {   // CTA-uniform per step
  int c1[2];                   // word index of site1 (i = 0, row 0) for sub-class A / B, lower half, class o
  int c2[2];                   // same for site2: upper half, class o'
  int rot1, rot2;              // row rotations (SPLIT: group A in the low 16 bits, group B in the high 16 bits)
  int si;                      // cyclic x shift of site2
  int flags;                   // bit0: x-parity of o, bit1: x-parity of o', bit2: site2 takes the other sub-class
};

template <int BX, int BY, int BZ, int MARGIN, int PXP, int PLP, bool SPLIT>
__device__ __forceinline__ void brw_make_dense_step(uint32_t k0, uint32_t k1, uint32_t step, uint32_t box_id,
                                                    uint32_t phase_lo, BrwDenseStep *out) {
  using G = BrwDenseGeom<BX, BY, BZ, MARGIN>;
  BrwPhilox4 r = brw_philox(0xFFFFFFFFu, step, box_id, phase_lo, k0, k1);
  // residue classes mod 4 of the bcc lattice: parity bit + three "half" bits; 16 classes each
  const uint32_t q1 = r.x & 15u, q2 = (r.x >> 4) & 15u;
  int o[2][3];
  o[0][0] = (q1 & 1) + 2 * ((q1 >> 1) & 1); o[0][1] = (q1 & 1) + 2 * ((q1 >> 2) & 1); o[0][2] = (q1 & 1) + 2 * ((q1 >> 3) & 1);
  o[1][0] = (q2 & 1) + 2 * ((q2 >> 1) & 1); o[1][1] = (q2 & 1) + 2 * ((q2 >> 2) & 1); o[1][2] = (q2 & 1) + 2 * ((q2 >> 3) & 1);
  for (int h = 0; h < 2; h++) {
    const int y_lo = h == 0 ? MARGIN : G::Y_UPPER;
    for (int sub = 0; sub < 2; sub++) {
      // first site of the sub-class inside the region: coordinate = lo + ((class + 2*sub - lo) mod 4)
      const int X = MARGIN + ((o[h][0] + 2 * sub - MARGIN) & 3);
      const int Y = y_lo + ((o[h][1] + 2 * sub - y_lo) & 3);
      const int Z = MARGIN + ((o[h][2] + 2 * sub - MARGIN) & 3);
      const int c = Z * PLP + (Y >> 1) * PXP + (X >> 1);
      if (h == 0) out->c1[sub] = c; else out->c2[sub] = c;
    }
  }
  if (SPLIT) {
    out->rot1 = (int)brw_below(r.y & 0xFFFF0000u, G::NRA) | (int)brw_below(r.y << 16, G::NRB) << 16;
    out->rot2 = (int)brw_below(r.z & 0xFFFF0000u, G::NRA) | (int)brw_below(r.z << 16, G::NRB) << 16;
  } else {
    out->rot1 = (int)brw_below(r.y, G::NROWS);
    out->rot2 = (int)brw_below(r.z, G::NROWS);
  }
  out->si = (int)(((r.w & 0xFFFFu) * (uint32_t)G::NI) >> 16);
  out->flags = (int)(q1 & 1u) | (int)((q2 & 1u) << 1) | (int)(((r.w >> 16) & 1u) << 2);
}

template <int LAT, int NSH, int PX, int PY, int PZ, int MARGIN, int PXP, int PLP, int NLIMB, bool EXACT, bool PAIRW,
          bool SPLIT = false>
__global__ void __launch_bounds__(1024) brw_box_metropolis_word_kernel(
    BrwGeom g, BrwBoxParams p, uint8_t *__restrict__ lat, const double *__restrict__ beta,
    const double *__restrict__ tab_g, const int4 *__restrict__ classes, const int4 *__restrict__ disp, uint32_t k0,
    uint32_t k1, uint32_t phase_lo, int mode, unsigned long long *__restrict__ att_out,
    unsigned long long *__restrict__ acc_out, double *__restrict__ dE_out) {
  static_assert(LAT == 1, "dense sets are derived for bcc");
  using T = BrwWordTab<NSH, NLIMB>;
  using G = BrwDenseGeom<2 * PX, 2 * PY, PZ, MARGIN>;
  (void)mode; (void)classes; (void)disp;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int *urow = reinterpret_cast<int *>(smem_raw);                              // [(PAIRS+1)][32][2]
  int *off = urow + T::urow_words;                                            // [2][ztot]
  const int tab_words = (T::urow_words + 2 * g.ztot + 1) & ~1;
  double *Vs = reinterpret_cast<double *>(urow + tab_words);                  // [n_shells][S][S]
  double *red = Vs + p.v_entries;                                             // [32]
  int *rowoff = reinterpret_cast<int *>(red + 32);                            // [2*NROWS]
  BrwDenseStep *sp = reinterpret_cast<BrwDenseStep *>(                        // [steps], 16-byte aligned
      smem_raw + (((size_t)tab_words * 4 + (size_t)p.v_entries * 8 + 32 * 8 + 2 * G::NROWS * 4 + 15) & ~(size_t)15));
  const int sp_steps = SPLIT ? max(p.steps, p.steps_a) : p.steps;
  uint32_t *wbox = reinterpret_cast<uint32_t *>(sp + sp_steps + 1);           // [bzc][PLP]
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

  {
    const int *tg = reinterpret_cast<const int *>(tab_g);
    for (int i = tid; i < tab_words + 2 * p.v_entries; i += blockDim.x) urow[i] = tg[i];
  }
  // rows of one sub-class in one half: r -> (j = r % NJ rows of 4 in y, pl = r / NJ planes of 4 in z); doubled so that
  // (warp + rotation) needs no wrap
  if (SPLIT) {
    // group A: rowoff[0 .. 2 NRA), planes [0, NPA); group B: rowoff[2 NRA .. 2 NRA + 2 NRB), planes [NPA + 1, NP)
    for (int r = tid; r < 2 * (G::NRA + G::NRB); r += blockDim.x) {
      const bool b = r >= 2 * G::NRA;
      const int rr = b ? (r - 2 * G::NRA) % G::NRB : r % G::NRA;
      rowoff[r] = 2 * PXP * (rr % G::NJ) + 4 * PLP * (rr / G::NJ + (b ? G::NPA + 1 : 0));
    }
  } else {
    for (int r = tid; r < 2 * G::NROWS; r += blockDim.x) {
      const int rr = r % G::NROWS;
      rowoff[r] = 2 * PXP * (rr % G::NJ) + 4 * PLP * (rr / G::NJ);
    }
  }
  for (int st = tid; st <= sp_steps; st += blockDim.x)
    brw_make_dense_step<2 * PX, 2 * PY, PZ, MARGIN, PXP, PLP, SPLIT>(k0, k1, (uint32_t)st, box_id, phase_lo, &sp[st]);
  brw_wbox_copy<LAT, PX, PY, PXP, PLP, false, PAIRW>(g, L, wbox, 0, PY * p.bzc, ox, oy, oz);
  __syncthreads();

  const double my_beta = beta[replica];
  const double beta_l2e = my_beta * 1.4426950408889634;
  // relative band of the fast acceptance test: ex2.approx + f32 argument rounding (< 1e-5 for |x| <= 126)
  // plus the dE guard propagated through exp
  const double band = my_beta * p.guard + 4e-5;
  const float bandf = (float)(my_beta * p.guard) + 1e-6f;
  const float c0 = (float)(beta_l2e * p.fix_scale), c20 = (float)(beta_l2e * p.fix_scale * 1048576.0);
  const long long gfix = (long long)ceil(p.guard / p.fix_scale) + 1;     // guard band in fixed-point units
  long long efix_sum = 0;                                                 // accepted fixed-point dE (exact)
  const int warp = tid >> 5, lane = tid & 31;
  const int sub = lane >= G::NI ? 1 : 0;
  const int li = lane - sub * G::NI;                     // position in the x-row
  const int grp = SPLIT && warp >= G::NRA ? 1 : 0;            // SPLIT: warp group (z zone)
  const int n_work = SPLIT ? G::NRA + G::NRB : G::NROWS;      // warps that hold a row pair
  const bool active = lane < 2 * G::NI && warp < n_work;
  const int *my_rowoff = rowoff + (SPLIT ? (grp ? 2 * G::NRA + (warp - G::NRA) : warp) : warp);
  const int row_mul = p.row_mul;
  unsigned int n_att = 0, n_acc = 0;
  double dE_sum = 0.0;
  BrwPhilox4 rnd = {0, 0, 0, 0};

  // One trial = a GATHER part (site words, Philox word, integer neighbour-count differences D[n] and the fixed-point
  // table row: shared-memory pipe) and a DECIDE part (dp4a dE, acceptance test, swap: ALU/FMA pipes).  The state
  // between the two lives in registers.
  uint32_t *w1 = wbox, *w2 = wbox;
  uint32_t wa = 0, wb = 0, rw = 0, D[NSH];
  int2 Tw[T::PAIRS + 1];
  int flags = 0;
  bool distinct = false;
  auto gather = [&](int step) {
    const BrwDenseStep q = sp[step];
    const int rot1 = SPLIT ? (grp ? q.rot1 >> 16 : q.rot1 & 0xFFFF) : q.rot1;
    const int rot2 = SPLIT ? (grp ? q.rot2 >> 16 : q.rot2 & 0xFFFF) : q.rot2;
    distinct = false;
    if (active) {
      int i2 = li + q.si; if (i2 >= G::NI) i2 -= G::NI;
      const int sub2 = sub ^ ((q.flags >> 2) & 1);
      w1 = wbox + (sub ? q.c1[1] : q.c1[0]) + 2 * li + my_rowoff[rot1];
      w2 = wbox + (sub2 ? q.c2[1] : q.c2[0]) + 2 * i2 + my_rowoff[rot2];
      flags = q.flags;
      wa = PAIRW ? (*w1 & 0xFFFFu) : *w1; wb = PAIRW ? (*w2 & 0xFFFFu) : *w2;
      n_att++;
      if ((step & 3) == 0) rnd = brw_philox((uint32_t)tid, (uint32_t)step, box_id, phase_lo, k0, k1);
      rw = (step & 3) == 0 ? rnd.x : (step & 3) == 1 ? rnd.y : (step & 3) == 2 ? rnd.z : rnd.w;
      distinct = wa != wb;
      if (distinct && !EXACT) {
        const int ca = PAIRW ? brw_pair_code_alu(wa) : brw_word_code(wa), cb = PAIRW ? brw_pair_code_alu(wb) : brw_word_code(wb);
        const int2 *row = reinterpret_cast<const int2 *>(urow) + (ca * row_mul + cb);
        if (BRW_TABG) {
#pragma unroll
          for (int j = 0; j <= T::PAIRS; j++) Tw[j] = row[j * 32];
        }
        uint32_t C1[NSH], C2[NSH];
        if (BRW_XP & 2) {
#pragma unroll
          for (int n = 0; n < NSH; n++) { C1[n] = wa * (n + 3); C2[n] = wb * (n + 5); }
        } else if (PAIRW) {
          if (flags & 1) BrwPairGatherBcc<NSH, PXP, PLP, 1>::run(w1, C1);
          else BrwPairGatherBcc<NSH, PXP, PLP, 0>::run(w1, C1);
          if (flags & 2) BrwPairGatherBcc<NSH, PXP, PLP, 1>::run(w2, C2);
          else BrwPairGatherBcc<NSH, PXP, PLP, 0>::run(w2, C2);
        } else {
          if (flags & 1) brw_wsum_shells<LAT, NSH, PXP, PLP, 1, 0>(w1, C1);
          else brw_wsum_shells<LAT, NSH, PXP, PLP, 0, 0>(w1, C1);
          if (flags & 2) brw_wsum_shells<LAT, NSH, PXP, PLP, 1, 0>(w2, C2);
          else brw_wsum_shells<LAT, NSH, PXP, PLP, 0, 0>(w2, C2);
        }
#pragma unroll
        for (int n = 0; n < NSH; n++) D[n] = 0x80808080u + C1[n] - C2[n];
      }
    }
  };
  auto decide = [&]() {
    if (active) {
      if (distinct) {
        bool decided = false, accept = false;
        double dE = 0.0;
        if (BRW_XP & 1) { decided = true; accept = ((rw ^ D[0] ^ D[1] ^ D[2] ^ D[3]) & 3u) == 0u; }
        else if (!EXACT) {
          if (!BRW_TABG) {
            const int ca = PAIRW ? brw_pair_code_alu(wa) : brw_word_code(wa), cb = PAIRW ? brw_pair_code_alu(wb) : brw_word_code(wb);
            const int2 *row = reinterpret_cast<const int2 *>(urow) + (ca * row_mul + cb);
#pragma unroll
            for (int j = 0; j <= T::PAIRS; j++) Tw[j] = row[j * 32];
          }
          int Sk[NLIMB];
#pragma unroll
          for (int k = 0; k < NLIMB; k++) Sk[k] = 0;
#pragma unroll
          for (int j = 0; j < T::PAIRS; j++) {
            const int e0 = 2 * j, e1 = 2 * j + 1;
            Sk[e0 % NLIMB] = brw_dp4a_us(D[e0 / NLIMB], Tw[j].x, Sk[e0 % NLIMB]);
            Sk[e1 % NLIMB] = brw_dp4a_us(D[e1 / NLIMB], Tw[j].y, Sk[e1 % NLIMB]);
          }
          const int2 Kw = Tw[T::PAIRS];
          long long efix = -(long long)(((unsigned long long)(uint32_t)Kw.y << 32) | (uint32_t)Kw.x);
#pragma unroll
          for (int k = 0; k < NLIMB; k++) efix += (long long)Sk[k] * (1LL << (8 * k));
          if (BRW_INTDEC) {
          // Decision on the FMA/ALU pipes only (no conversions, no MUFU, no fp64: nothing that goes through the MIO
            // queue the other warp group's gathers are filling).  dE = efix * 2^-k exactly; outside the guard band its
            // sign is the reference's.  For dE > 0: x = -beta*log2(e)*dE in f32 from two 20-bit chunks (magic-number
            // int->float), t = 2^x by round-to-nearest split + degree-7 polynomial + exponent add, u from the top 23 bits
            // of the Philox word.  |t/t_true - 1| < 3e-7 + 1.4e-7 |x| (+ beta*guard), |u_f - u| < 2^-23: the trial is decided
            // here only if |u_f - t| exceeds three times those bounds, else by the reference association below.
            if (efix < -gfix) { accept = true; decided = true; }
            else if (efix > gfix) {
              const uint32_t lo = (uint32_t)efix & 0xFFFFFu, hi = min((uint32_t)(efix >> 20), 0x7FFFFFu);
              const float flo = __int_as_float(0x4B000000u | lo) - 8388608.0f;
              const float fhi = __int_as_float(0x4B000000u | hi) - 8388608.0f;
              const float x = fmaxf(-fmaf(fhi, c20, flo * c0), -100.0f);
              const float xr = x + 12582912.0f;
              const float f = x - (xr - 12582912.0f);
              const float y = f * 0.693147180559945f;
              float pz = 1.98412698e-4f;
              pz = fmaf(pz, y, 1.38888889e-3f); pz = fmaf(pz, y, 8.33333333e-3f); pz = fmaf(pz, y, 4.16666667e-2f);
              pz = fmaf(pz, y, 1.66666667e-1f); pz = fmaf(pz, y, 0.5f); pz = fmaf(pz, y, 1.0f); pz = fmaf(pz, y, 1.0f);
              const int nexp = __float_as_int(xr) - 0x4B400000;
              const float t = __int_as_float(__float_as_int(pz) + (nexp << 23));
              const float uf = __int_as_float(0x3F800000u | (rw >> 9)) - 1.0f;
              if (fabsf(uf - t) > fmaf(t, fmaf(fabsf(x), 5e-7f, bandf), 2.5e-7f)) { accept = uf < t; decided = true; }
            }
          } else {
            // fixed-point sign test, then ex2.approx.ftz.f32 (relative error < 2^-22, argument rounded to f32:
            // < 1e-5 for |x| <= 126) with a correspondingly wide band
            if (efix < -gfix) { accept = true; decided = true; }
            else if (efix > gfix) {
              const double t = (double)brw_ex2_approx((float)(-beta_l2e * ((double)efix * p.fix_scale)));
              const double u = brw_u01(rw);
              if (fabs(u - t) > t * band) { accept = u < t; decided = true; }
            }
          }
          if (decided && accept) efix_sum += efix;
        }
        if (!decided) {
          // reference association, generic loop (screened kernel: ~1e-5 of the trials)
          const double u = brw_u01(rw);
          const int ca = PAIRW ? brw_pair_code(wa) : brw_word_code(wa), cb = PAIRW ? brw_pair_code(wb) : brw_word_code(wb);
          const int sa = brw_code_species(ca), sb = brw_code_species(cb);
          const int S = g.S;
          const int *off1 = off + (flags & 1) * g.ztot, *off2 = off + ((flags >> 1) & 1) * g.ztot;
          double E1a = 0.0, E1b = 0.0, E2b = 0.0, E2a = 0.0;
          int k = 0;
#pragma unroll 1
          for (int n = 0; n < NSH; n++) {
            double e1a = 0.0, e1b = 0.0, e2b = 0.0, e2a = 0.0;
            const double *Vn = Vs + n * S * S;
            const int end = g.shell_end[n];
#pragma unroll 1
            for (; k < end; k++) {
              const int s1 = brw_code_species(PAIRW ? brw_pair_code(brw_lo16(w1 + off1[k])) : brw_word_code(w1[off1[k]]));
              const int s2 = brw_code_species(PAIRW ? brw_pair_code(brw_lo16(w2 + off2[k])) : brw_word_code(w2[off2[k]]));
              e1a = __dadd_rn(e1a, Vn[s1 * S + sa]); e1b = __dadd_rn(e1b, Vn[s1 * S + sb]);
              e2b = __dadd_rn(e2b, Vn[s2 * S + sb]); e2a = __dadd_rn(e2a, Vn[s2 * S + sa]);
            }
            if (n == 0) { E1a = e1a; E1b = e1b; E2b = e2b; E2a = e2a; }
            else { E1a = __dadd_rn(E1a, e1a); E1b = __dadd_rn(E1b, e1b); E2b = __dadd_rn(E2b, e2b); E2a = __dadd_rn(E2a, e2a); }
          }
          const double before = __dadd_rn(E1a, E2b);             // pair_energy, sites unswapped
          const double after = __dadd_rn(E1b, E2a);              // pair_energy, sites swapped
          dE = __dsub_rn(after, before);                         // src/metropolis.F90:792
          accept = dE < 0.0;                                     // :796
          if (!accept) accept = u < exp(-my_beta * dE);          // :802
          if (accept) dE_sum += dE;
        }
        if (accept) {
          if (PAIRW) {
            // a site's nibbles live in the low half of its own word and in the high half of its left neighbour's
            uint16_t *h1 = reinterpret_cast<uint16_t *>(w1), *h2 = reinterpret_cast<uint16_t *>(w2);
            h1[0] = (uint16_t)wb; h1[-1] = (uint16_t)wb;
            h2[0] = (uint16_t)wa; h2[-1] = (uint16_t)wa;
          } else { *w1 = wb; *w2 = wa; }
          n_acc++;
        }
      } else n_acc++;                                            // :774-777
    }
  };

  // the groups are independent, and group A (fewer warps) finishes a step sooner: it runs steps_a >= steps steps so
  // that both groups end the phase together
  const int my_steps = SPLIT && !grp && !BRW_ANTI ? p.steps_a : p.steps;
  if (SPLIT && BRW_ANTI) {
    // experiment: the two groups in enforced anti-phase, one CTA barrier per half step (group A: gather s | decide s;
    // group B: decide s-1 | gather s), so that the shared-memory-pipe half of one group runs beside the ALU half of
    // the other.  Needs a decide half without MIO-queue instructions (BRW_INTDEC, BRW_TABG) to have a chance.
    const int halves = 2 * p.steps + 1;
    for (int hs = 0; hs < halves; hs++) {
      const int t = hs - grp;
      if (warp < n_work && t >= 0 && t < 2 * p.steps) {
        if (t & 1) decide(); else gather(t >> 1);
      }
      __syncthreads();
    }
  } else
  for (int step = 0; step < my_steps; step++) {
    if (SPLIT && warp >= n_work) break;                        // no row pair: not a member of either group barrier
    gather(step);
    decide();
    if (SPLIT) {
      // the two groups never touch each other's active sites: independent barriers (the groups drift apart, a slow
      // trial -- reference-association path -- stalls only its own group)
      if (grp) asm volatile("bar.sync 2, %0;" ::"n"(32 * G::NRB) : "memory");
      else asm volatile("bar.sync 1, %0;" ::"n"(32 * G::NRA) : "memory");
    } else __syncthreads();
  }
  if (SPLIT) __syncthreads();

  // frozen margin planes are unchanged (and, with a z pitch < PZ, shared with the neighbouring box): not stored
  brw_wbox_copy<LAT, PX, PY, PXP, PLP, true, PAIRW>(g, L, wbox, PY * MARGIN, PY * (p.bzc - MARGIN), ox, oy, oz);
  dE_sum += (double)efix_sum * p.fix_scale;
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
