// epoch_metropolis.cuh -- production Metropolis (nbr_energy, pair_energy,
// pair_swap, monte_carlo_step_lattice; src/metropolis.F90:751-813), pair-word lattice + dense non-interacting sets
// (word_metropolis.cuh) with the SITE ENERGIES CACHED over an EPOCH of K steps.
//
// Observation.  In a step of the dense decomposition every touched site belongs to S_o (lower-y half of the box) or to
// S_o' (upper-y half), and no two of them interact.  A swap changes only the species ON those sites, so the
// neighbour-count vector C_i[shell][species] of every site of S_o u S_o' -- which counts the species AROUND the site,
// not on it -- stays valid for as long as the two classes (o, o') and the row assignment are kept, and with it the
// energy e_i(a) = sum_n sum_s C_i[n][s] V_n(a, s) of putting ANY species a on site i.  Because the two sites of a trial
// do not interact,  dE(a on 1 <-> b on 2) = [e_1(b) + e_2(a)] - [e_1(a) + e_2(b)].  So, per epoch:
//
//   epoch (CTA-uniform o, o', row rotations):  every lane gathers C of its home site of S_o and of its home site of
//       S_o' (2 x 30 shared-memory loads + nibble arithmetic, pair_gather.inc) and turns each into the fixed-point
//       vector h_i[a] = e_i(a) - e_i(reference species: 4, or 3 when S <= 4) (dp4a against three signed 8-bit digits of
//       the table, which sits in the kernel-parameter constant bank), stored in a per-warp cache [site kind][c][slot];
//   K steps (CTA-uniform x shift and A<->B exchange per step):  lane l proposes  home site 1 (l)  <->  home site 2 of
//       lane perm_j(l) of the SAME warp.  A trial is five conflict-free LDS.32 (the partner's species and four cache
//       entries), three integer adds, and an f32 acceptance test (ex2.approx); any trial whose fixed-point dE or
//       acceptance test lies inside the rigorous error band is recomputed with the reference's f64 association
//       (src/bw_hamiltonian.f90:171-173, :1014-1017, :111-112; src/metropolis.F90:792-802) and decided by it, so every
//       accept/reject is the one the reference arithmetic yields (EXACT = true runs that path for every trial; the
//       tests compare the two instantiations trajectory for trajectory).
//
// The partner of a lane is always a site of its own warp's row pair, so the K steps of an epoch need only __syncwarp;
// the CTA-wide barrier is per EPOCH (gathers of the next epoch read what other warps wrote) and split-phase -- an mbarrier
// in shared memory: a warp arrives after its last step, prepares the next epoch's parameters / addresses / Philox draw,
// then waits -- so the warps drift apart inside an epoch and their shared-memory and ALU phases overlap.  One launch is
// one phase; consecutive phases are programmatic dependent launches (griddepcontrol), and the box arrives through the
// bulk-async copy engine (cp.async.bulk: brw_pbox_load_tma).  Every step is still a set of pairwise non-interacting swap
// proposals whose choice does not depend on the configuration, each its own inverse: detailed balance holds move by
// move exactly as before, and the simultaneous decisions equal sequential ones.  What K costs is sampling efficiency:
// a site is tried K times in a row against an unchanged neighbourhood (measured in DESIGN.md 4.3: relaxation per
// attempted swap relative to the reference's sequential sampler).
//
// Geometry: boxes are 68 x 64 x 32 doubled-grid units at a PITCH of 64 x 64 x 28: consecutive boxes share their frozen
// margin planes in x and z, which leaves 60 active units = 15 sites per class row (30 of 32 lanes; 28 in the 64-wide
// box).  32 warps take 32 of the 36 (A row, B row) pairs per half and epoch (rotation: all rows are visited).
#pragma once
#include "word_metropolis.cuh"

template <int K> struct __align__(16) BrwEpochT {   // CTA-uniform per epoch
  int c1[2];                   // word index of home site 1 (i = 0, row 0) for sub-class A / B: lower half, class o
  int c2[2];                   // same for home site 2: upper half, class o'
  int rot;                     // row rotations: site-1 rows in the low 16 bits, site-2 rows in the high 16 bits
  int flags;                   // bit0: x-parity of o, bit1: x-parity of o'
  uint32_t sw[2];              // per step j: byte j = cyclic shift of the 2*NI slots (A row, then B row) of the row pair
};

template <int BX, int BY, int BZ, int MARGIN, int PXP, int PLP, int K>
__device__ __forceinline__ void brw_make_epoch(uint32_t k0, uint32_t k1, uint32_t epoch, uint32_t box_id,
                                               uint32_t phase_lo, BrwEpochT<K> *out) {
  using G = BrwDenseGeom<BX, BY, BZ, MARGIN>;
  static_assert(K <= 8, "eight step bytes per epoch");
  const BrwPhilox4 r = brw_philox(0xFFFFFFFFu, 2u * epoch, box_id, phase_lo, k0, k1);
  const BrwPhilox4 t = brw_philox(0xFFFFFFFFu, 2u * epoch + 1u, box_id, phase_lo, k0, k1);
  // residue classes mod 4 of the bcc lattice: parity bit + three "half" bits; 16 classes each
  const uint32_t q1 = r.x & 15u, q2 = (r.x >> 4) & 15u;
  int o[2][3];
  o[0][0] = (q1 & 1) + 2 * ((q1 >> 1) & 1); o[0][1] = (q1 & 1) + 2 * ((q1 >> 2) & 1); o[0][2] = (q1 & 1) + 2 * ((q1 >> 3) & 1);
  o[1][0] = (q2 & 1) + 2 * ((q2 >> 1) & 1); o[1][1] = (q2 & 1) + 2 * ((q2 >> 2) & 1); o[1][2] = (q2 & 1) + 2 * ((q2 >> 3) & 1);
  for (int h = 0; h < 2; h++) {
    const int y_lo = h == 0 ? MARGIN : G::Y_UPPER;
    for (int sub = 0; sub < 2; sub++) {
      const int X = MARGIN + ((o[h][0] + 2 * sub - MARGIN) & 3);
      const int Y = y_lo + ((o[h][1] + 2 * sub - y_lo) & 3);
      const int Z = MARGIN + ((o[h][2] + 2 * sub - MARGIN) & 3);
      const int c = Z * PLP + (Y >> 1) * PXP + (X >> 1);
      if (h == 0) out->c1[sub] = c; else out->c2[sub] = c;
    }
  }
  out->rot = (int)brw_below(r.y, G::NROWS) | (int)brw_below(r.z, G::NROWS) << 16;
  out->flags = (int)(q1 & 1u) | (int)((q2 & 1u) << 1);
  const uint32_t rs[4] = {r.w, t.x, t.y, t.z};
  uint32_t sw[2] = {0u, 0u};
  for (int j = 0; j < K; j++) {
    const uint32_t h16 = (rs[j >> 1] >> (16 * (j & 1))) & 0xFFFFu;
    const uint32_t b = (h16 * (uint32_t)(2 * G::NI)) >> 16;
    sw[j >> 2] |= b << (8 * (j & 3));
  }
  out->sw[0] = sw[0]; out->sw[1] = sw[1];
}

// box <-> global copy for boxes of PX = 32 + XT compact sites per row, one warp per row (lane = x; lanes < XT also
// take x = 32 + lane).  STORE = false: global bytes -> pair words; STORE = true: low halves of the pair words -> global
// bytes, without the frozen x margins (shared with the neighbouring boxes: read by both, written by nobody).
template <int PX, int PY, int PXP, int PLP, int XM, bool STORE>
__device__ __forceinline__ void brw_pbox_copy(const BrwGeom &g, uint8_t *L, uint32_t *wbox, int row_begin, int row_end,
                                              int ox, int oy, int oz) {
  constexpr int XT = PX - 32;
  static_assert(XT >= 0 && XT < 32, "a row is 32..63 sites");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  int gx0 = (ox >> 1) + lane; if (gx0 >= g.cx) gx0 -= g.cx; if (gx0 >= g.cx) gx0 -= g.cx;
  int gx1 = gx0 + 32; if (gx1 >= g.cx) gx1 -= g.cx; if (gx1 >= g.cx) gx1 -= g.cx;
  // CTAs of PY warps: warp w always copies compact-y row w of a plane, the plane index advances by one per iteration
  const bool row_per_warp = nwarps == PY && row_begin % PY == 0;
#pragma unroll 8
  for (int r = row_begin + warp; r < row_end; r += nwarps) {
    const int lyc = row_per_warp ? warp : r % PY, lz = row_per_warp ? (r - warp) / PY : r / PY;
    int gzz = oz + lz; if (gzz >= g.gz) gzz -= g.gz; if (gzz >= g.gz) gzz -= g.gz;
    int gyy = oy + 2 * lyc + (lz & 1); if (gyy >= g.gy) gyy -= g.gy; if (gyy >= g.gy) gyy -= g.gy;
    uint8_t *grow = L + ((long)gzz * g.cy + (gyy >> 1)) * g.cx;
    uint32_t *w = wbox + lz * PLP + lyc * PXP;
    if (STORE) {
      if (lane >= XM) grow[gx0] = (uint8_t)brw_code_species(brw_pair_code(w[lane]));
      if (lane < XT - XM) grow[gx1] = (uint8_t)brw_code_species(brw_pair_code(w[32 + lane]));
    } else {
      const uint32_t v0 = brw_species_nibbles(grow[gx0]);
      const uint32_t v1 = lane < XT ? brw_species_nibbles(grow[gx1]) : 0u;
      uint32_t h0 = __shfl_down_sync(0xffffffffu, v0, 1);
      const uint32_t h1 = __shfl_down_sync(0xffffffffu, v1, 1);       // lane XT-1 receives 0: beyond the row, never a neighbour
      const uint32_t first1 = __shfl_sync(0xffffffffu, v1, 0);
      if (lane == 31) h0 = first1;
      w[lane] = v0 | (h0 << 16);
      if (lane < XT) w[32 + lane] = v1 | (h1 << 16);
    }
  }
}

__device__ __forceinline__ int brw_lds32(uint32_t a) { int v; asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void brw_sts32(uint32_t a, int v) { asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// split-phase CTA barrier (mbarrier in shared memory): a warp ARRIVES when its epoch is done, prepares what the next epoch
// needs that no other warp can change (epoch parameters, addresses, the Philox draw), then WAITS for the other warps
#ifndef BRW_SPLITBAR
#define BRW_SPLITBAR 1
#endif
#ifndef BRW_MBAR_SLEEP
#define BRW_MBAR_SLEEP 0
#endif
#ifndef BRW_MBAR_HINT
#define BRW_MBAR_HINT 20000u
#endif
__device__ __forceinline__ void brw_mbar_init(uint32_t a, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void brw_mbar_arrive(uint32_t a) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
}
__device__ __forceinline__ void brw_mbar_wait(uint32_t a, uint32_t parity) {
  uint32_t ok;
  do {
    // suspend-time hint (ns): the warp sleeps in hardware until the phase completes instead of re-issuing the poll (the
    // polls of waiting warps were 7 % of the executed instructions, taking issue slots from the warps still working)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(a), "r"(parity), "r"(BRW_MBAR_HINT) : "memory");
#if BRW_MBAR_SLEEP
    if (!ok) __nanosleep(BRW_MBAR_SLEEP);                  // waiting warps stay out of the issue slots of the working ones
#endif
  } while (!ok);
}

// ---- box load through the bulk-async copy engine (TMA, cp.async.bulk) -----------------------------------------------------
// Warp w owns compact-y row w of every plane of the box.  Each lane asks the copy engine for the whole global x-row (g.cx
// bytes <= 128: 16-byte aligned and a multiple of 16, checked by the host) of one plane, straight INTO the storage of the
// pair-word row it will become (136 bytes; the bytes go to its first 16-byte aligned address), all on one mbarrier per
// warp (transaction bytes).  All 1024 row copies of a box are in flight at once and no register holds data in flight;
// when they have landed the warp expands its rows in place -- every lane reads its bytes, then the words are written over
// them.  The periodic wrap in y and z is resolved per row by the source address, the wrap in x by the index into the row.
__device__ __forceinline__ void brw_mbar_expect_tx(uint32_t a, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void brw_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
template <int PX, int PY, int PZ, int PXP, int PLP>
__device__ __forceinline__ void brw_pbox_load_tma(const BrwGeom &g, const uint8_t *L, uint32_t *wbox, unsigned long long *mbars,
                                                  int ox, int oy, int oz) {
  constexpr int XT = PX - 32;
  static_assert(PXP * 4 >= 128 + 8 && (PLP * 4) % 16 == 0, "a 128-byte row fits behind the alignment gap of a pair-word row");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;           // blockDim = 32 * PY: warp = compact-y row of a plane
  const uint32_t cx = (uint32_t)g.cx;
  const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(mbars + warp);
  const uint32_t row0 = (uint32_t)__cvta_generic_to_shared(wbox + warp * PXP);      // this warp's row of plane 0
  const uint32_t gap = (row0 & 8u);                                                 // PXP * 4 = 8 mod 16, PLP * 4 = 0 mod 16
  int gx0 = (ox >> 1) + lane; if (gx0 >= g.cx) gx0 -= g.cx; if (gx0 >= g.cx) gx0 -= g.cx;
  int gx1 = gx0 + 32; if (gx1 >= g.cx) gx1 -= g.cx; if (gx1 >= g.cx) gx1 -= g.cx;
  if (lane == 0) {
    brw_mbar_init(mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    brw_mbar_expect_tx(mbar, (uint32_t)PZ * cx);
  }
  __syncwarp();
  for (int lz = lane; lz < PZ; lz += 32) {
    int gzz = oz + lz; if (gzz >= g.gz) gzz -= g.gz; if (gzz >= g.gz) gzz -= g.gz;
    int gyy = oy + 2 * warp + (lz & 1); if (gyy >= g.gy) gyy -= g.gy; if (gyy >= g.gy) gyy -= g.gy;
    brw_bulk_g2s(row0 + (uint32_t)lz * (PLP * 4) + gap, L + ((long)gzz * g.cy + (gyy >> 1)) * g.cx, cx, mbar);
  }
  brw_mbar_wait(mbar, 0u);
  const uint8_t *bytes = reinterpret_cast<const uint8_t *>(wbox + warp * PXP) + gap;
  constexpr int CH = 8;
  static_assert(PZ % CH == 0, "planes in chunks of eight");
#pragma unroll 1
  for (int base = 0; base < PZ; base += CH) {
    uint32_t v0[CH], v1[CH];
#pragma unroll
    for (int k = 0; k < CH; k++) {
      const uint8_t *srow = bytes + (size_t)(base + k) * (PLP * 4);
      v0[k] = brw_species_nibbles(srow[gx0]);
      v1[k] = lane < XT ? brw_species_nibbles(srow[gx1]) : 0u;
    }
    __syncwarp();                                                          // every lane has read its bytes of these rows
#pragma unroll
    for (int k = 0; k < CH; k++) {
      uint32_t h0 = __shfl_down_sync(0xffffffffu, v0[k], 1);
      const uint32_t h1 = __shfl_down_sync(0xffffffffu, v1[k], 1);       // lane XT-1 receives 0: beyond the row, never a neighbour
      const uint32_t first1 = __shfl_sync(0xffffffffu, v1[k], 0);
      if (lane == 31) h0 = first1;
      uint32_t *w = wbox + (base + k) * PLP + warp * PXP;
      w[lane] = v0[k] | (h0 << 16);
      if (lane < XT) w[32 + lane] = v1[k] | (h1 << 16);
    }
  }
}

// signed 8-bit digits of the fixed-point site-energy table (kernel parameter: read straight from the constant bank)
#define BRW_HLIMB 3
#ifndef BRW_EXP
#define BRW_EXP 0       // timing experiment only (results invalid): bit 0 never takes the reference-association path
#endif

// The decision of a trial that the cached fixed-point energies cannot take with certainty (~1e-5 of the trials).  Inlined:
// as a __noinline__ call it was 6 % slower (call ABI in the step loop); a build that never takes this path at all (invalid
// results) is 8 % faster still -- tools/xp_run.py.
//  1. second screening level: dE in f64 from the exact integer neighbour counts of the two sites,
//     dE = sum_n sum_s (c1 - c2)[n][s] (V_n(b, s) - V_n(a, s)) -- the reference's value up to f64 rounding (~1e-16 Z max|V|).
//     Outside its own guard band (guard2 = 1e-9 Z max|V|, and the same band propagated through exp) the decision is the
//     reference's;
//  2. inside (~1e-9 of the trials), or always when EXACT: the reference association (src/bw_hamiltonian.f90:171-173,
//     :1014-1017, :111-112; src/metropolis.F90:792-802), generic loop.
template <int NSH, int PXP, int PLP, bool EXACT>
__device__ __forceinline__ bool brw_epoch_decide_cold(const uint32_t *w1, const uint32_t *w2, int flags, uint32_t wa, uint32_t wb,
                                                   uint32_t rw, double my_beta, const double *Vs, const int *off, int S, int ztot,
                                                   int4 shell_end, double guard2, double *dE_out) {
  const double u = brw_u01(rw);
  const int sa = brw_code_species(brw_pair_code(wa)), sb2 = brw_code_species(brw_pair_code(wb));
  bool accept = false;
  double dE = 0.0;
  if (!EXACT) {
    uint32_t C1[NSH], C2[NSH];
    if (flags & 1) BrwPairGatherBcc<NSH, PXP, PLP, 1>::run(w1, C1); else BrwPairGatherBcc<NSH, PXP, PLP, 0>::run(w1, C1);
    if (flags & 2) BrwPairGatherBcc<NSH, PXP, PLP, 1>::run(w2, C2); else BrwPairGatherBcc<NSH, PXP, PLP, 0>::run(w2, C2);
#pragma unroll
    for (int n = 0; n < NSH; n++) {
      const double *Vn = Vs + n * S * S;
      int rest = 0;
#pragma unroll
      for (int s2 = 0; s2 < 4; s2++) {
        if (s2 < S) {
          const int d = (int)((C1[n] >> (8 * s2)) & 255u) - (int)((C2[n] >> (8 * s2)) & 255u);
          rest -= d;
          dE = fma((double)d, Vn[s2 * S + sb2] - Vn[s2 * S + sa], dE);
        }
      }
      if (S == 5) dE = fma((double)rest, Vn[4 * S + sb2] - Vn[4 * S + sa], dE);
    }
    if (fabs(dE) > guard2) {
      if (dE < 0.0) { *dE_out = dE; return true; }
      const double t = exp(-my_beta * dE);
      if (fabs(u - t) > t * (my_beta * guard2 + 1e-12)) { *dE_out = dE; return u < t; }
    }
  }
  const int *off1 = off + (flags & 1) * ztot, *off2 = off + ((flags >> 1) & 1) * ztot;
  const int ends[4] = {shell_end.x, shell_end.y, shell_end.z, shell_end.w};
  double E1a = 0.0, E1b = 0.0, E2b = 0.0, E2a = 0.0;
  int k = 0;
#pragma unroll 1
  for (int n = 0; n < NSH; n++) {
    double e1a = 0.0, e1b = 0.0, e2b = 0.0, e2a = 0.0;
    const double *Vn = Vs + n * S * S;
    const int end = n == 0 ? ends[0] : n == 1 ? ends[1] : n == 2 ? ends[2] : ends[3];
#pragma unroll 1
    for (; k < end; k++) {
      const int s1 = brw_code_species(brw_pair_code(brw_lo16(w1 + off1[k])));
      const int s2 = brw_code_species(brw_pair_code(brw_lo16(w2 + off2[k])));
      e1a = __dadd_rn(e1a, Vn[s1 * S + sa]); e1b = __dadd_rn(e1b, Vn[s1 * S + sb2]);
      e2b = __dadd_rn(e2b, Vn[s2 * S + sb2]); e2a = __dadd_rn(e2a, Vn[s2 * S + sa]);
    }
    if (n == 0) { E1a = e1a; E1b = e1b; E2b = e2b; E2a = e2a; }
    else { E1a = __dadd_rn(E1a, e1a); E1b = __dadd_rn(E1b, e1b); E2b = __dadd_rn(E2b, e2b); E2a = __dadd_rn(E2a, e2a); }
  }
  const double before = __dadd_rn(E1a, E2b);             // pair_energy, sites unswapped
  const double after = __dadd_rn(E1b, E2a);              // pair_energy, sites swapped
  dE = __dsub_rn(after, before);                         // src/metropolis.F90:792
  accept = dE < 0.0;                                     // :796
  if (!accept) accept = u < exp(-my_beta * dE);          // :802
  *dE_out = dE;
  return accept;
}

template <int NSH, int PX, int PY, int PZ, int MARGIN, int PXP, int PLP, int NLIMB, bool EXACT, int K>
__global__ void __launch_bounds__(1024) brw_box_metropolis_epoch_kernel(
    BrwGeom g, BrwBoxParams p, uint8_t *__restrict__ lat, const double *__restrict__ beta,
    const double *__restrict__ tab_g, const int4 *__restrict__ classes, const int4 *__restrict__ disp, uint32_t k0,
    uint32_t k1, uint32_t phase_lo, int mode, unsigned long long *__restrict__ att_out,
    unsigned long long *__restrict__ acc_out, double *__restrict__ dE_out) {
  using T = BrwWordTab<NSH, NLIMB>;
  using G = BrwDenseGeom<2 * PX, 2 * PY, PZ, MARGIN>;
  using Ep = BrwEpochT<K>;
  static_assert(K == 2 || K == 4 || K == 8, "one Philox call serves four steps");
  static_assert(2 * G::NI <= 32 && (PXP & 1) == 0, "A row on even, B row on odd word addresses");
  static_assert(NSH == 4, "gather plan and table layout are generated for four shells");
  (void)mode; (void)classes; (void)disp;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int *urow = reinterpret_cast<int *>(smem_raw);                              // [(PAIRS+1)][32][2] (unused here; blob layout kept)
  int *off = urow + T::urow_words;                                            // [2][ztot]
  const int tab_words = (T::urow_words + 2 * g.ztot + 1) & ~1;
  double *Vs = reinterpret_cast<double *>(urow + tab_words);                  // [n_shells][S][S]
  double *red = Vs + p.v_entries;                                             // [32]
  int *rowoff = reinterpret_cast<int *>(red + 32);                            // [2*NROWS]
  const int n_epochs = p.steps / K;
  Ep *ep = reinterpret_cast<Ep *>(                                            // [n_epochs], 16-byte aligned
      smem_raw + (((size_t)tab_words * 4 + (size_t)p.v_entries * 8 + 32 * 8 + 2 * G::NROWS * 4 + 15) & ~(size_t)15));
  // per-warp site-energy cache: h[site kind 0/1][row 0..4][slot 0..31], row = species + 1, row 0 (species 4) = zeros
  int *hcache = reinterpret_cast<int *>(ep + n_epochs);                       // [32 warps][2][5][32]
  uint32_t *wbox = reinterpret_cast<uint32_t *>(hcache + 32 * 320);           // [PZ][PLP]
  __shared__ unsigned int s_att[32], s_acc[32];
  __shared__ __align__(8) unsigned long long s_mbar, s_tma_bar[32];

  const int tid = threadIdx.x;
  const int replica = blockIdx.x / p.boxes_per_replica;
  const int bid = blockIdx.x - replica * p.boxes_per_replica;
  const int bi = bid % p.nb[0], bj = (bid / p.nb[0]) % p.nb[1], bk = bid / (p.nb[0] * p.nb[1]);
  uint8_t *L = lat + (long)replica * g.n_sites;
  const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(&s_mbar);
  if (BRW_SPLITBAR && tid == 0) brw_mbar_init(mbar, blockDim.x >> 5);       // one arrival per warp and epoch

  BrwPhilox4 ro = brw_philox(0xFFFFFFFEu, 0u, (uint32_t)replica, phase_lo, k0, k1);
  const int ox = 2 * (int)brw_below(ro.x, g.gx >> 1) + bi * p.B[0];
  const int oy = 2 * (int)brw_below(ro.y, g.gy >> 1) + bj * p.B[1];
  const int oz = 2 * (int)brw_below(ro.z, g.gz >> 1) + bk * p.B[2];
  const uint32_t box_id = (uint32_t)blockIdx.x;

  {
    const int *tg = reinterpret_cast<const int *>(tab_g);
    for (int i = tid; i < tab_words + 2 * p.v_entries; i += blockDim.x) urow[i] = tg[i];
  }
  for (int r = tid; r < 2 * G::NROWS; r += blockDim.x) {
    const int rr = r % G::NROWS;
    rowoff[r] = 2 * PXP * (rr % G::NJ) + 4 * PLP * (rr / G::NJ);
  }
  for (int e = tid; e < n_epochs; e += blockDim.x)
    brw_make_epoch<2 * PX, 2 * PY, PZ, MARGIN, PXP, PLP, K>(k0, k1, (uint32_t)e, box_id, phase_lo, &ep[e]);
  for (int i = tid; i < 32 * 2 * 32; i += blockDim.x) {
    hcache[(i >> 5) * 160 + (i & 31)] = 0;                                                         // row 0 = species 4
    hcache[(i >> 5) * 160 + 128 + (i & 31)] = 0;                                                   // row 4 = species 3: the reference species when S <= 4
  }
  // programmatic dependent launch: everything above reads constant tables only and may overlap the tail of the previous
  // phase's grid (a CTA of this grid starts as soon as an SM frees up); the lattice is read after the previous grid has
  // completed and flushed.  The next phase may be scheduled from now on (it waits at the same point).
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;");
  if (p.tma_stages > 0) brw_pbox_load_tma<PX, PY, PZ, PXP, PLP>(g, L, wbox, s_tma_bar, ox, oy, oz);
  else brw_pbox_copy<PX, PY, PXP, PLP, MARGIN / 2, false>(g, L, wbox, 0, PY * PZ, ox, oy, oz);
  __syncthreads();

  // fast acceptance test in f32: t = ex2.approx(x), x = fl(fl(efix) * c0): three roundings, |dx| <= 1.8e-7 |x|.  For |x| <= 32
  // that is < 4e-6 relative in t, plus 2^-22 of ex2.approx: band 6e-6 t (+ the dE guard propagated through exp); for
  // |x| > 32, t < 2.4e-10 and any relative error of t disappears in the absolute term 2.5e-7, which also covers u taken
  // from the top 23 bits of the Philox word (< 1.2e-7) and the rounding of u - t (< 6e-8)
  const float c0 = (float)(-beta[replica] * 1.4426950408889634 * p.fix_scale);
  const float bandf = (float)(beta[replica] * p.guard) + 6e-6f;
  const int gfix = p.gfix;                                                // guard band in fixed-point units
  long long efix_sum = 0;                                                 // accepted fixed-point dE (exact)
  const int warp = tid >> 5, lane = tid & 31;
  const bool active = lane < 2 * G::NI;
  // byte address (shared window) of this lane's column of the per-warp cache: row of species code c at + 128 c
  const uint32_t hb1 = (uint32_t)__cvta_generic_to_shared(hcache + warp * 320 + lane), hb2w = hb1 + 640 - 4 * lane;
  unsigned int n_acc = 0;
  BrwPhilox4 rnd = {0, 0, 0, 0};
  if (tid < 32) red[tid] = 0.0;                                           // sum of accepted dE of the reference-association path
  __syncthreads();

  // the part of an epoch's set-up that does not read the lattice: done for epoch e + 1 between the arrival at the epoch
  // barrier and the wait (BRW_SPLITBAR), so that it overlaps the wait for the slowest warp
  Ep E = ep[0];
  uint32_t *w1 = wbox, *row2 = wbox;
  const uint32_t *w2h = wbox;
  auto prepare = [&](int e) {
    E = ep[e];
    if (active) {
      const int sub = lane >= G::NI ? 1 : 0;
      row2 = wbox + rowoff[warp + (E.rot >> 16)];
      w1 = wbox + (sub ? E.c1[1] - 2 * G::NI : E.c1[0]) + 2 * lane + rowoff[warp + (E.rot & 0xFFFF)];
      w2h = row2 + (sub ? E.c2[1] - 2 * G::NI : E.c2[0]) + 2 * lane;
      if (K >= 4 || (e & 1) == 0) rnd = brw_philox((uint32_t)tid, (uint32_t)((e * K) >> 2), box_id, phase_lo, k0, k1);
    }
  };
  prepare(0);

  for (int e = 0; e < n_epochs; e++) {
    const int flags = E.flags;
    uint32_t wa = 0, ra = 0;
    // home sites: slot = lane; slots [0, NI) = sub-class A (even word addresses), [NI, 2 NI) = sub-class B (odd)
    const int c2A = E.c2[0], c2B = E.c2[1] - 2 * G::NI;
    const uint32_t sw0 = E.sw[0], sw1 = E.sw[1];
    if (active) {
      wa = *w1 & 0xFFFFu;
      ra = ((wa * 0x1234u) >> 5) & 0x380u;
      if (!EXACT) {
        // neighbour counts of both home sites (four byte fields per shell), then the fixed-point energy of every
        // species on the site relative to species 4: h[c] = sum_n sum_s C[n][s] * X_n[c][s], three signed digits
#pragma unroll
        for (int kind = 0; kind < 2; kind++) {
          uint32_t C[NSH];
          const uint32_t *wc = kind ? w2h : w1;
          if (flags & (1 << kind)) BrwPairGatherBcc<NSH, PXP, PLP, 1>::run(wc, C);
          else BrwPairGatherBcc<NSH, PXP, PLP, 0>::run(wc, C);
          const uint32_t hk = kind ? hb1 + 640 : hb1;
#pragma unroll
          for (int c = 1; c <= 4; c++) {
            if (c == 4 && p.h_rows < 4) break;                   // <= 4 species: energies relative to species 3, row 4 = zeros
            int sl[BRW_HLIMB];
#pragma unroll
            for (int l = 0; l < BRW_HLIMB; l++) {
              sl[l] = 0;
#pragma unroll
              for (int n = 0; n < NSH; n++) sl[l] = brw_dp4a_us(C[n], p.xdig[(n * BRW_HLIMB + l) * 4 + c - 1], sl[l]);
            }
            brw_sts32(hk + 128 * c, sl[0] + (sl[1] << 8) + (sl[2] << 16));
          }
        }
      }
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < K; j++) {
      const int gs = K >= 4 ? j : ((e & 1) * 2 + j);            // step index mod 4 (compile-time for K = 4, 8)
      if (active) {
        // partner = home site 2 of slot (lane + shift) mod 2 NI of this warp's row pair
        int slot2 = lane + (int)(((j >> 2 ? sw1 : sw0) >> (8 * (j & 3))) & 0xFFu);
        if (slot2 >= 2 * G::NI) slot2 -= 2 * G::NI;
        uint32_t *w2 = row2 + (slot2 >= G::NI ? c2B : c2A) + 2 * slot2;
        const uint32_t wb = *w2 & 0xFFFFu;
        if ((gs & 3) == 0 && j > 0) rnd = brw_philox((uint32_t)tid, (uint32_t)((e * K + j) >> 2), box_id, phase_lo, k0, k1);
        const uint32_t rw = (gs & 3) == 0 ? rnd.x : (gs & 3) == 1 ? rnd.y : (gs & 3) == 2 ? rnd.z : rnd.w;
        n_acc += wa == wb;                                          // :774-777
        if (wa != wb) {
          bool accept = false, fast = !EXACT;
          int efix = 0;
          const uint32_t rb = ((wb * 0x1234u) >> 5) & 0x380u;      // 128 * (species + 1), 0 for species 4
          if (!EXACT) {
            const uint32_t hb2 = hb2w + 4 * slot2;
            // dE = [e1(b) + e2(a)] - [e1(a) + e2(b)]: the two sites do not interact
            efix = (brw_lds32(hb1 + rb) - brw_lds32(hb2 + rb)) - (brw_lds32(hb1 + ra) - brw_lds32(hb2 + ra));
            if (efix < -gfix) accept = true;
            else if (efix > gfix) {
              const float t = brw_ex2_approx((float)efix * c0);
              const float d = (__int_as_float(0x3F800000u | (rw >> 9)) - 1.0f) - t;
              accept = d < 0.0f;
              fast = fabsf(d) > fmaf(t, bandf, 2.5e-7f);
            } else fast = false;
          }
          if ((BRW_EXP & 1) && !EXACT) fast = true;
          if (!fast) {
            // ~1e-5 of the trials: second screening level in f64, then the reference association
            efix = 0;
            double dE;
            accept = brw_epoch_decide_cold<NSH, PXP, PLP, EXACT>(w1, w2, flags, wa, wb, rw, beta[replica], Vs, off, g.S, g.ztot,
                                                                 make_int4(g.shell_end[0], g.shell_end[1], g.shell_end[2], g.shell_end[3]),
                                                                 p.guard2, &dE);
            if (accept) atomicAdd(&red[warp], dE);
          }
          if (accept) {
            // a site's nibbles live in the low half of its own word and in the high half of its left neighbour's
            uint16_t *q1 = reinterpret_cast<uint16_t *>(w1), *q2 = reinterpret_cast<uint16_t *>(w2);
            q1[0] = (uint16_t)wb; q1[-1] = (uint16_t)wb;
            q2[0] = (uint16_t)wa; q2[-1] = (uint16_t)wa;
            wa = wb; ra = rb;
            n_acc++;
            efix_sum += efix;
          }
        }
      }
      __syncwarp();
    }
    if (BRW_SPLITBAR) {
      if (lane == 0) brw_mbar_arrive(mbar);                  // the __syncwarp above ordered the warp's stores before it
      if (e + 1 < n_epochs) prepare(e + 1);
      brw_mbar_wait(mbar, (uint32_t)(e & 1));
    } else {
      __syncthreads();
      if (e + 1 < n_epochs) prepare(e + 1);
    }
  }

  // frozen margin planes are unchanged (and shared with the neighbouring box in z): not stored
  brw_pbox_copy<PX, PY, PXP, PLP, MARGIN / 2, true>(g, L, wbox, PY * MARGIN, PY * (PZ - MARGIN), ox, oy, oz);
  // every active lane attempts one trial per step
  unsigned int n_att = active ? (unsigned int)(n_epochs * K) : 0u;
  double dE_sum = (double)efix_sum * p.fix_scale;
  for (int o = 16; o > 0; o >>= 1) {
    n_att += __shfl_down_sync(0xffffffffu, n_att, o);
    n_acc += __shfl_down_sync(0xffffffffu, n_acc, o);
    dE_sum += __shfl_down_sync(0xffffffffu, dE_sum, o);
  }
  if ((tid & 31) == 0) { s_att[tid >> 5] = n_att; s_acc[tid >> 5] = n_acc; red[tid >> 5] += dE_sum; }
  __syncthreads();
  if (tid == 0) {
    unsigned long long A = 0, C = 0; double D = 0.0;
    for (int w = 0; w < (int)((blockDim.x + 31) >> 5); w++) { A += s_att[w]; C += s_acc[w]; D += red[w]; }
    att_out[blockIdx.x] += A; acc_out[blockIdx.x] += C; dE_out[blockIdx.x] += D;
  }
}
