// epoch_byte_metropolis.cuh -- the epoch scheme of epoch_metropolis.cuh (site energies cached over K steps, warp-local
// pairing, one CTA barrier per epoch) on the BYTE lattice with the period-P decomposition of tile_metropolis.cuh, for
// the geometries that have no dense-set word kernel: fcc (4 / 6 shells) and bcc with 6 shells (SURVEY 8a rows a4/a5/
// a9/a10; src/metropolis.F90:751-813; neighbour tables src/bw_hamiltonian.f90:440-509, 1310-1683).
//
// Epoch (CTA-uniform residue class o and displacement class d, drawn like a step of brw_box_metropolis_fast_kernel):
// thread t = coarse cell (i,j,k) owns home site 1 = m + o + P (i,j,k) and home site 2 = m + (o+d mod P) + P (i,j,k).
// The planner guarantees that ALL sites of the two classes inside the active region are pairwise non-interacting, so
// (a) their neighbour counts -- compile-time gathers brw_count_shells: LDS.U8 + shl + add per neighbour -- and hence the
// fixed-point energy vectors h[a] = e(a) - e(reference species) stay valid for the whole epoch, and (b) any pairing of
// a home site 1 with a home site 2 is a valid Metropolis trial.  Step: lane l proposes its home site 1 <-> the home site
// 2 of lane (l + shift) mod n_lanes of its own warp (CTA-uniform shift): five shared-memory loads, three adds, an f32
// acceptance pre-test; any trial inside the rigorous error band is recomputed with the reference's f64 association
// (brw_fast_shells: sequential per shell, shells left to right) and decided by it.  EXACT = true takes that path for every
// trial on the same schedule (the trajectory-identity test compares the two instantiations).
#pragma once
#include "epoch_metropolis.cuh"

template <int K> struct __align__(16) BrwByteEpochT {
  BrwStepParams q;             // c1_base, c2_base, par1, par2 of the epoch's (o, d); q.s unused
  uint32_t sh[2];              // byte j: the shift of step j as a fraction of 256 of the warp's lane count
  int rot;                     // rotation of the thread -> coarse cell map (so the 32 cells that share a warp change)
  uint32_t pad;
};

// reference association for one trial (~1e-9 of the trials of the screened kernel): one out-of-line copy per instantiation,
// not one per unrolled step -- inlined, the unrolled chains bloat the step loop (measured: -20 %)
template <int LAT, int NSH, int PX, int PY>
__device__ __noinline__ double brw_byte_exact_dE(const uint8_t *box, const char *Vl, int S, int c1, int c2, int par1, int par2,
                                                 int sa, int sb) {
  double E1a, E1b, E2b, E2a;
  if (par1) brw_fast_shells<LAT, NSH, PX, PY, 1, 0>(box + c1, Vl, S, sa, sb, E1a, E1b);
  else brw_fast_shells<LAT, NSH, PX, PY, 0, 0>(box + c1, Vl, S, sa, sb, E1a, E1b);
  if (par2) brw_fast_shells<LAT, NSH, PX, PY, 1, 0>(box + c2, Vl, S, sb, sa, E2b, E2a);
  else brw_fast_shells<LAT, NSH, PX, PY, 0, 0>(box + c2, Vl, S, sb, sa, E2b, E2a);
  const double before = __dadd_rn(E1a, E2b);           // pair_energy, sites unswapped
  const double after = __dadd_rn(E1b, E2a);            // pair_energy, sites swapped
  return __dsub_rn(after, before);                     // src/metropolis.F90:792
}

// The decision of a trial that the cached fixed-point energies cannot take with certainty (~1e-5 of the trials).  1. second screening level: f64 dE from the exact integer counts of both
// sites; outside its guard band (guard2 = 1e-9 Z max|V|, propagated through exp) the decision is the reference's.
// 2. inside, or always when EXACT: the reference association (brw_fast_shells: sequential per shell, shells left to right).
template <int LAT, int NSH, int PX, int PY, bool EXACT>
__device__ __forceinline__ bool brw_byte_decide_cold(const uint8_t *box, const char *Vl, int S, int c1, int c2, int par1, int par2,
                                                  int sa, int sb, uint32_t rw, double my_beta, double guard2, double *dE_out) {
  const double u = brw_u01(rw);
  double dE = 0.0;
  if (!EXACT) {
    const uint32_t box_s = (uint32_t)__cvta_generic_to_shared(box);
    uint32_t C1[NSH], C2[NSH];
    if (par1) brw_count_shells<LAT, NSH, PX, PY, 1, 0>(box_s + c1, C1); else brw_count_shells<LAT, NSH, PX, PY, 0, 0>(box_s + c1, C1);
    if (par2) brw_count_shells<LAT, NSH, PX, PY, 1, 0>(box_s + c2, C2); else brw_count_shells<LAT, NSH, PX, PY, 0, 0>(box_s + c2, C2);
#pragma unroll
    for (int n = 0; n < NSH; n++) {
      const double *Va = reinterpret_cast<const double *>(Vl + ((n * S + sa) * S) * 128);
      const double *Vb = reinterpret_cast<const double *>(Vl + ((n * S + sb) * S) * 128);
      int rest = 0;
#pragma unroll
      for (int sp = 0; sp < 4; sp++) {
        if (sp < S) {
          const int d = (int)((C1[n] >> (8 * sp)) & 255u) - (int)((C2[n] >> (8 * sp)) & 255u);
          rest -= d;
          dE = fma((double)d, Vb[sp * 16] - Va[sp * 16], dE);
        }
      }
      if (S == 5) dE = fma((double)rest, Vb[4 * 16] - Va[4 * 16], dE);
    }
    if (fabs(dE) > guard2) {
      if (dE < 0.0) { *dE_out = dE; return true; }
      const double t = exp(-my_beta * dE);
      if (fabs(u - t) > t * (my_beta * guard2 + 1e-12)) { *dE_out = dE; return u < t; }
    }
  }
  dE = brw_byte_exact_dE<LAT, NSH, PX, PY>(box, Vl, S, c1, c2, par1, par2, sa, sb);
  bool accept = dE < 0.0;                              // :796
  if (!accept) accept = u < exp(-my_beta * dE);        // :802
  *dE_out = dE;
  return accept;
}

// PX = compact sites per x-row of the box, PITCH >= PX = bytes between consecutive rows in shared memory: with PITCH = PX =
// 32 the 2-3 rows of coarse cells a warp covers start 128 bytes apart (fcc) and their gathers collide in the banks (43 %
// replays measured); the padded pitches below were chosen with a bank simulation of all period orientations.
template <int LAT, int NSH> struct BrwBytePitch { static constexpr int value = LAT == 2 && NSH <= 4 ? 56 : 52; };

template <int LAT, int NSH, int PX, int PY, bool EXACT, int K, int MAXT>
__global__ void __launch_bounds__(MAXT) brw_box_metropolis_byte_epoch_kernel(
    BrwGeom g, BrwBoxParams p, uint8_t *__restrict__ lat, const double *__restrict__ beta,
    const double *__restrict__ Vrep, const int4 *__restrict__ classes, const int4 *__restrict__ disp, uint32_t k0,
    uint32_t k1, uint32_t phase_lo, int mode, unsigned long long *__restrict__ att_out,
    unsigned long long *__restrict__ acc_out, double *__restrict__ dE_out) {
  using Ep = BrwByteEpochT<K>;
  constexpr int PITCH = BrwBytePitch<LAT, NSH>::value;
  static_assert(PITCH >= PX && PITCH % 4 == 0, "padded row pitch");
  static_assert(K == 4 || K == 8, "one Philox call serves four steps");
  static_assert(NSH * BRW_HLIMB * 4 <= 72, "xdig table of BrwBoxParams");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const BrwBoxMode &md = p.mode[mode];
  double *Vs = reinterpret_cast<double *>(smem_raw);                       // [v_entries][16] lane-replicated V
  double *red = Vs + p.v_entries * 16;                                     // [32]
  const int n_epochs = p.steps / K;
  Ep *ep = reinterpret_cast<Ep *>(red + 32);                               // [n_epochs]
  const int n_warps = (MAXT + 31) / 32;
  int *hcache = reinterpret_cast<int *>(ep + n_epochs);                    // [n_warps][2][5][32]; rows = species, zeros for the reference
  int *base_tab = hcache + n_warps * 320;                                  // [MAXT] compact box offset of every thread's coarse cell
  uint8_t *box = reinterpret_cast<uint8_t *>(base_tab + MAXT);             // [bzc][PY][PITCH], bytes hold 8*species
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
  for (int e = tid; e < n_epochs; e += blockDim.x) {
    brw_make_step<0>(g, p, md, classes, disp, k0, k1, (uint32_t)e, box_id, phase_lo, &ep[e].q);
    ep[e].q.c1_base = (ep[e].q.c1_base / PX) * PITCH + ep[e].q.c1_base % PX;       // row pitch PX (p.bxc) -> PITCH
    ep[e].q.c2_base = (ep[e].q.c2_base / PX) * PITCH + ep[e].q.c2_base % PX;
    const BrwPhilox4 t = brw_philox(0xFFFFFFFCu, (uint32_t)e, box_id, phase_lo, k0, k1);
    ep[e].sh[0] = t.x; ep[e].sh[1] = t.y;
    ep[e].rot = (int)brw_below(t.z, (uint32_t)md.M);
  }
  for (int i = tid; i < n_warps * 320; i += blockDim.x) hcache[i] = 0;     // the reference species' rows stay zero
  const int stx = md.P[0] >> 1, sty = (LAT == 1 ? (md.P[1] >> 1) : md.P[1]) * PITCH, stz = md.P[2] * PY * PITCH;
  const bool active = tid < md.M;
  {
    const int A0 = md.A[0], A1 = md.A[1];
    const int ci = tid % A0, cr = tid / A0, cj = cr % A1, ck = cr / A1;
    if (tid < MAXT) base_tab[tid] = ci * stx + cj * sty + ck * stz;
  }
  // programmatic dependent launch (see epoch_metropolis.cuh): the table set-up above overlaps the previous phase's tail
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;");
  brw_box_copy<LAT, PX, PY, false, PITCH>(g, L, box, PY * p.bzc, ox, oy, oz);
  if (tid < 32) red[tid] = 0.0;
  __syncthreads();

  const char *Vl = reinterpret_cast<const char *>(Vs + (tid & 15));
  const uint32_t box_s = (uint32_t)__cvta_generic_to_shared(box);
  const int S = g.S;
  const float c0 = (float)(-beta[replica] * 1.4426950408889634 * p.fix_scale);
  const float bandf = (float)(beta[replica] * p.guard) + 6e-6f;
  const int gfix = p.gfix;
  const int warp = tid >> 5, lane = tid & 31;
  const int n_lanes = min(32, md.M - 32 * warp);                          // coarse cells held by this warp (<= 0: none)
  const uint32_t hb1 = (uint32_t)__cvta_generic_to_shared(hcache + warp * 320 + lane), hb2w = hb1 + 640 - 4 * lane;
  long long efix_sum = 0;
  unsigned int n_acc = 0;
  BrwPhilox4 rnd = {0, 0, 0, 0};

  // split-phase epoch barrier (see epoch_metropolis.cuh): a warp arrives when its epoch is done, prepares what the next epoch
  // needs that no other warp can change (epoch parameters, its cell, the Philox draw), then waits for the other warps
  __shared__ __align__(8) unsigned long long s_mbar;
  const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(&s_mbar);
  if (BRW_SPLITBAR && tid == 0) brw_mbar_init(mbar, blockDim.x >> 5);
  Ep E = ep[0];
  int cell0 = 0, base1 = 0;
  auto prepare = [&](int e) {
    E = ep[e];
    // this epoch's coarse cell of the thread: cell = (tid + rot) mod M; a warp holds n_lanes consecutive cells (mod M)
    cell0 = 32 * warp + E.rot;
    if (cell0 >= md.M) cell0 -= md.M;
    base1 = 0;
    if (active) {
      int cell = cell0 + lane; if (cell >= md.M) cell -= md.M;
      base1 = base_tab[cell];
      rnd = brw_philox((uint32_t)tid, (uint32_t)((e * K) >> 2), box_id, phase_lo, k0, k1);
    }
  };
  prepare(0);
  __syncthreads();                                                         // the mbarrier is initialised

  for (int e = 0; e < n_epochs; e++) {
    const int c1 = E.q.c1_base + base1;
    int a = 0;
    if (active) {
      a = box[c1];
      if (!EXACT) {
#pragma unroll
        for (int kind = 0; kind < 2; kind++) {
          uint32_t C[NSH];
          const uint32_t sa = box_s + (kind ? E.q.c2_base + base1 : c1);
          if (kind ? E.q.par2 : E.q.par1) brw_count_shells<LAT, NSH, PITCH, PY, 1, 0>(sa, C);
          else brw_count_shells<LAT, NSH, PITCH, PY, 0, 0>(sa, C);
          const uint32_t hk = kind ? hb1 + 640 : hb1;
#pragma unroll
          for (int c = 0; c < 4; c++) {
            if (c == 3 && p.h_rows < 4) break;                             // <= 4 species: relative to species 3 (row 3 = zeros)
            int sl[BRW_HLIMB];
#pragma unroll
            for (int l = 0; l < BRW_HLIMB; l++) {
              sl[l] = 0;
#pragma unroll
              for (int n = 0; n < NSH; n++) sl[l] = brw_dp4a_us(C[n], p.xdig[(n * BRW_HLIMB + l) * 4 + c], sl[l]);
            }
            brw_sts32(hk + 128 * c, sl[0] + (sl[1] << 8) + (sl[2] << 16));
          }
        }
      }
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < K; j++) {
      if (active) {
        int slot2 = lane + (int)((((E.sh[j >> 2] >> (8 * (j & 3))) & 0xFFu) * (uint32_t)n_lanes) >> 8);
        if (slot2 >= n_lanes) slot2 -= n_lanes;
        int cell2 = cell0 + slot2;
        if (cell2 >= md.M) cell2 -= md.M;
        const int c2 = E.q.c2_base + base_tab[cell2];
        const int b = box[c2];
        if ((j & 3) == 0 && j > 0) rnd = brw_philox((uint32_t)tid, (uint32_t)((e * K + j) >> 2), box_id, phase_lo, k0, k1);
        const uint32_t rw = (j & 3) == 0 ? rnd.x : (j & 3) == 1 ? rnd.y : (j & 3) == 2 ? rnd.z : rnd.w;
        n_acc += a == b;                                                   // :774-777
        if (a != b) {
          bool accept = false, fast = !EXACT;
          int efix = 0;
          if (!EXACT) {
            // box bytes hold 8*species: row offset = 128 * species = byte << 4
            const uint32_t ra = (uint32_t)a << 4, rb = (uint32_t)b << 4, hb2 = hb2w + 4 * slot2;
            efix = (brw_lds32(hb1 + rb) - brw_lds32(hb2 + rb)) - (brw_lds32(hb1 + ra) - brw_lds32(hb2 + ra));
            if (efix < -gfix) accept = true;
            else if (efix > gfix) {
              const float t = brw_ex2_approx((float)efix * c0);
              const float d = (__int_as_float(0x3F800000u | (rw >> 9)) - 1.0f) - t;
              accept = d < 0.0f;
              fast = fabsf(d) > fmaf(t, bandf, 2.5e-7f);
            } else fast = false;
          }
          if (!fast) {
            // ~1e-5 of the trials: second screening level in f64, then the reference association
            efix = 0;
            double dE;
            accept = brw_byte_decide_cold<LAT, NSH, PITCH, PY, EXACT>(box, Vl, S, c1, c2, E.q.par1, E.q.par2, a >> 3, b >> 3, rw,
                                                                   beta[replica], p.guard2, &dE);
            if (accept) atomicAdd(&red[warp], dE);
          }
          if (accept) {
            box[c1] = (uint8_t)b; box[c2] = (uint8_t)a;
            a = b;
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

  brw_box_copy<LAT, PX, PY, true, PITCH>(g, L, box, PY * p.bzc, ox, oy, oz);
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
