// word_metropolis.cuh -- production Metropolis (nbr_energy,
// pair_energy, pair_swap, monte_carlo_step_lattice; src/metropolis.F90:751-813) on a WORD lattice.
//
// Same decomposition, same Philox counters and the same decisions as brw_box_metropolis_fast_kernel
// (tile_metropolis.cuh) -- trajectories are identical -- but the box is held in shared memory as one
// 32-bit word per site, word = 1 << (8*species) for species 0..3 and 0 for species 4.  The integer
// neighbour counts of the screened dE then need ONE LDS.32 + half an IADD3 per neighbour:
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
// a correspondingly wider band.
//
// Rows are padded (pitch PXP words = 32 + A0, plane pitch PLP) so that the 32 lanes of a warp --
// consecutive coarse cells, 3 words apart in x -- fall into 32 different banks: bank = 3*lane.
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

// box <-> global copy, one warp per compact-x row of PX sites.  STORE=false: global bytes -> shared words.
template <int LAT, int PX, int PY, int PXP, int PLP, bool STORE>
__device__ __forceinline__ void brw_wbox_copy(const BrwGeom &g, uint8_t *L, uint32_t *wbox, int n_rows, int ox, int oy,
                                              int oz) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  static_assert(PX <= 32, "one warp covers a row");
  const int oxc = ox >> 1;
#pragma unroll 8
  for (int r = warp; r < n_rows; r += nwarps) {
    const int lyc = r % PY, lz = r / PY;
    int gzz = oz + lz; if (gzz >= g.gz) gzz -= g.gz; if (gzz >= g.gz) gzz -= g.gz;
    const int Y = LAT == 1 ? 2 * lyc + (lz & 1) : lyc;
    int gyy = oy + Y; if (gyy >= g.gy) gyy -= g.gy; if (gyy >= g.gy) gyy -= g.gy;
    if (lane < PX) {
      int gxc = oxc + lane; if (gxc >= g.cx) gxc -= g.cx; if (gxc >= g.cx) gxc -= g.cx;
      const long gi = ((long)gzz * g.cy + (gyy >> g.ys)) * g.cx + gxc;
      uint32_t *w = wbox + lz * PLP + lyc * PXP + lane;
      if (STORE) L[gi] = (uint8_t)brw_code_species(brw_word_code(*w));
      else *w = brw_species_word(L[gi]);
    }
  }
}

// Constant tables of the word kernel, one blob in global memory copied to shared memory per CTA:
//   int32  urow[25][ROWP]   row (code_a*5 + code_b): words [n*NLIMB + k] = the k-th signed digit of U*2^k for the four
//                           species fields packed as int8x4; words [NSH*NLIMB .. +1] = int64 K = 128 * sum_f Ufix_f
//   int32  off[2][ztot]     word offsets of the neighbours per x-parity (fallback path)
//   double V[n_shells][S][S] reference V_ex in Fortran order V(centre, nbr, shell) (fallback path)
template <int NSH, int NLIMB> struct BrwWordTab {
  static constexpr int ROWP = NSH * NLIMB + 4;      // 20 (4 limbs) / 28 (6 limbs): 8 rows start in 8 different bank quads
  static constexpr int urow_words = 25 * ROWP;
};

// P0..P2 / A0..A2: the (single) period orientation and the coarse-cell counts of the plan, compile-time so
// that the per-step index arithmetic uses immediates (the host checks them against the plan).
template <int LAT, int NSH, int PX, int PY, int PXP, int PLP, int NLIMB, int MAXT, int P0, int P1, int P2, int A0, int A1,
          int A2>
__global__ void __launch_bounds__(MAXT) brw_box_metropolis_word_kernel(
    BrwGeom g, BrwBoxParams p, uint8_t *__restrict__ lat, const double *__restrict__ beta,
    const double *__restrict__ tab_g, const int4 *__restrict__ classes, const int4 *__restrict__ disp, uint32_t k0,
    uint32_t k1, uint32_t phase_lo, int mode, unsigned long long *__restrict__ att_out,
    unsigned long long *__restrict__ acc_out, double *__restrict__ dE_out) {
  using T = BrwWordTab<NSH, NLIMB>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const BrwBoxMode &md = p.mode[0];                                           // one orientation (P0,P1,P2)
  (void)mode;
  int *urow = reinterpret_cast<int *>(smem_raw);                              // [25][ROWP]
  int *off = urow + T::urow_words;                                            // [2][ztot]
  const int tab_words = (T::urow_words + 2 * g.ztot + 1) & ~1;
  double *Vs = reinterpret_cast<double *>(urow + tab_words);                  // [n_shells][S][S]
  double *red = Vs + p.v_entries;                                             // [32]
  BrwStepParams *sp = reinterpret_cast<BrwStepParams *>(                      // [steps], 16-byte aligned
      smem_raw + (((size_t)tab_words * 4 + (size_t)p.v_entries * 8 + 32 * 8 + 15) & ~(size_t)15));
  uint32_t *wbox = reinterpret_cast<uint32_t *>(sp + p.steps);                // [bzc][PLP]
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
  for (int st = tid; st < p.steps; st += blockDim.x) {
    BrwStepParams q;
    brw_make_step<0>(g, p, md, classes, disp, k0, k1, (uint32_t)st, box_id, phase_lo, &q);
    // re-express the two base sites in the padded word layout (make_step used pitch bxc/byc)
    const int z1 = q.c1_base / (PX * PY), r1 = q.c1_base - z1 * PX * PY;
    const int z2 = q.c2_base / (PX * PY), r2 = q.c2_base - z2 * PX * PY;
    q.c1_base = z1 * PLP + (r1 / PX) * PXP + (r1 % PX);
    q.c2_base = z2 * PLP + (r2 / PX) * PXP + (r2 % PX);
    sp[st] = q;
  }
  brw_wbox_copy<LAT, PX, PY, PXP, PLP, false>(g, L, wbox, PY * p.bzc, ox, oy, oz);
  __syncthreads();

  const double my_beta = beta[replica];
  const double beta_l2e = my_beta * 1.4426950408889634;
  // relative band of the fast acceptance test: ex2.approx + f32 argument rounding (< 1e-5 for |x| <= 126)
  // plus the dE guard propagated through exp
  const double band = my_beta * p.guard + 4e-5;
  constexpr int stx = P0 >> 1, sty = (LAT == 1 ? (P1 >> 1) : P1) * PXP, stz = P2 * PLP;
  const bool active = tid < A0 * A1 * A2;
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
      uint32_t *w1 = wbox + q.c1_base + base1;
      uint32_t *w2 = wbox + q.c2_base + i2 * stx + j2 * sty + k2 * stz;
      const uint32_t wa = *w1, wb = *w2;
      n_att++;
      if ((step & 3) == 0) rnd = brw_philox((uint32_t)tid, (uint32_t)step, box_id, phase_lo, k0, k1);
      if (wa != wb) {
        const uint32_t w = (step & 3) == 0 ? rnd.x : (step & 3) == 1 ? rnd.y : (step & 3) == 2 ? rnd.z : rnd.w;
        const double u = brw_u01(w);
        uint32_t C1[NSH], C2[NSH];
        if (q.par1) brw_wsum_shells<LAT, NSH, PXP, PLP, 1, 0>(w1, C1);
        else brw_wsum_shells<LAT, NSH, PXP, PLP, 0, 0>(w1, C1);
        if (q.par2) brw_wsum_shells<LAT, NSH, PXP, PLP, 1, 0>(w2, C2);
        else brw_wsum_shells<LAT, NSH, PXP, PLP, 0, 0>(w2, C2);
        const int ca = brw_word_code(wa), cb = brw_word_code(wb);
        const int4 *row = reinterpret_cast<const int4 *>(urow + (ca * 5 + cb) * T::ROWP);
        int Sk[NLIMB];
#pragma unroll
        for (int k = 0; k < NLIMB; k++) Sk[k] = 0;
#pragma unroll
        for (int j = 0; j < NSH * NLIMB / 4; j++) {
          const int4 Lw = row[j];
          const int e0 = 4 * j, e1 = 4 * j + 1, e2 = 4 * j + 2, e3 = 4 * j + 3;
          Sk[e0 % NLIMB] = brw_dp4a_us(0x80808080u + C1[e0 / NLIMB] - C2[e0 / NLIMB], Lw.x, Sk[e0 % NLIMB]);
          Sk[e1 % NLIMB] = brw_dp4a_us(0x80808080u + C1[e1 / NLIMB] - C2[e1 / NLIMB], Lw.y, Sk[e1 % NLIMB]);
          Sk[e2 % NLIMB] = brw_dp4a_us(0x80808080u + C1[e2 / NLIMB] - C2[e2 / NLIMB], Lw.z, Sk[e2 % NLIMB]);
          Sk[e3 % NLIMB] = brw_dp4a_us(0x80808080u + C1[e3 / NLIMB] - C2[e3 / NLIMB], Lw.w, Sk[e3 % NLIMB]);
        }
        long long efix = -*reinterpret_cast<const long long *>(urow + (ca * 5 + cb) * T::ROWP + NSH * NLIMB);
#pragma unroll
        for (int k = 0; k < NLIMB; k++) efix += (long long)Sk[k] * (1LL << (8 * k));
        double dE = (double)efix * p.fix_scale;                  // exact: |efix| < 2^53, fix_scale = 2^-k
        bool decided = false, accept = false;
        if (fabs(dE) > p.guard) {
          if (dE < 0.0) { accept = true; decided = true; }
          else {
            const double t = (double)brw_ex2_approx((float)(-beta_l2e * dE));
            if (fabs(u - t) > t * band) { accept = u < t; decided = true; }
          }
        }
        if (!decided) {
          // reference association, generic loop (rare: ~1e-5 of trials)
          const int sa = brw_code_species(ca), sb = brw_code_species(cb);
          const int S = g.S;
          double E1a = 0.0, E1b = 0.0, E2b = 0.0, E2a = 0.0;
          int k = 0;
#pragma unroll 1
          for (int n = 0; n < NSH; n++) {
            double e1a = 0.0, e1b = 0.0, e2b = 0.0, e2a = 0.0;
            const double *Vn = Vs + n * S * S;
            const int end = g.shell_end[n];
#pragma unroll 1
            for (; k < end; k++) {
              const int s1 = brw_code_species(brw_word_code(w1[off[q.par1 * g.ztot + k]]));
              const int s2 = brw_code_species(brw_word_code(w2[off[q.par2 * g.ztot + k]]));
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
        }
        if (accept) { *w1 = wb; *w2 = wa; n_acc++; dE_sum += dE; }
      } else n_acc++;                                            // :774-777
    }
    __syncthreads();
  }

  brw_wbox_copy<LAT, PX, PY, PXP, PLP, true>(g, L, wbox, PY * p.bzc, ox, oy, oz);
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
