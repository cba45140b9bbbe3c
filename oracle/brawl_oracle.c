/*
 * brawl_oracle.c -- CPU restatement of BraWl's atom-swap Monte-Carlo hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (brawl_b200/, libbrawl_cuda.so)
 * may include, link or call this file; it is the checker used by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs.
 *
 * Parity status: PINNED.  This restatement reproduces the reference's own golden files
 * (tests/99_ref/01_serial_metropolis, 02_parallel_metropolis ranks 0-3,
 * 03_serial_nested_sampling) bit-for-bit -- see tests/test_oracle_golden.py, which checks it
 * against fixtures built from those files by tests/golden/make_golden.py.  The MT19937
 * restated here is checked against the reference's own src/mt19937ar.c compiled into
 * oracle/_ref/ (oracle/Makefile) whenever /root/reference is present.
 * Exception -- parity UNPINNED: the Wang-Landau control-plane arithmetic at the end of this file
 * (orc_wl_mean_energy, orc_wl_window_optimise).  The reference ships no golden output for its load
 * balancer; these two are a second, independent restatement against which the host mirror in
 * brawl_b200/wang_landau.py is compared.  The WL trial loop itself (orc_wl_sweeps) is pinned
 * statistically by tests/99_ref/04_parallel_wang-landau/wl_dos.nc (1 % NRMSE, the reference's criterion).
 *
 * Each function cites the reference file:line it follows (paths relative to
 * /root/reference).  Conventions: all coordinates are 0-based here (the reference is
 * 1-based); the species grid is int8 grid[z][y][x] with x fastest == Fortran
 * config(1,x,y,z) column-major (src/shared_data.f90:30, src/initialise.F90:295); 0 = no
 * lattice site.  V is the raw Fortran-ordered array V_ex(centre, nbr, shell)
 * (src/io.f90:401): V[(shell*S + nbr)*S + centre] with 0-based species.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "shell_tables.inc"

#define ORC_SC 0
#define ORC_BCC 1
#define ORC_FCC 2

/* ------------------------------------------------------------------------------------------
 * MT19937 (Matsumoto & Nishimura 2002), restated from the published algorithm; the reference
 * vendors it as src/mt19937ar.c:63-125 with global state.  State is explicit here.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  uint32_t mt[624];
  int32_t mti;
} orc_mt;

void orc_mt_init(orc_mt *g, uint32_t seed) { /* src/mt19937ar.c:63-78 */
  g->mt[0] = seed;
  for (int i = 1; i < 624; i++)
    g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
  g->mti = 624;
}

/* src/mt19937ar.c:127-142 static-seed branch: seed = 110179 + 11*rank */
uint32_t orc_mt_static_seed(int rank) { return (uint32_t)(110179 + 11 * rank); }

uint32_t orc_mt_int32(orc_mt *g) { /* src/mt19937ar.c:82-118 */
  if (g->mti >= 624) {
    uint32_t *mt = g->mt;
    for (int k = 0; k < 624; k++) {
      uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
      mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    g->mti = 0;
  }
  uint32_t y = g->mt[g->mti++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

double orc_mt_genrand(orc_mt *g) { /* src/mt19937ar.c:121-125 */
  return orc_mt_int32(g) * (1.0 / 4294967296.0);
}

/* ------------------------------------------------------------------------------------------
 * System description
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int lattice, n1, n2, n3, S, n_shells;
  int gx, gy, gz;       /* grid extents 2*n (src/initialise.F90:295) */
  int wx, wy, wz;       /* neighbour wrap moduli: 2n, except simple cubic wraps with n
                           (src/bw_hamiltonian.f90:1944-1949 -- reference quirk) */
  int n_atoms;          /* src/initialise.F90:162,175,214 */
  int z_total;          /* neighbours per site over the n_shells shells */
  int shell_start[10], shell_count[10];
  const signed char (*off)[3];
  double *V;
} orc_sys;

orc_sys *orc_sys_create(int lattice, int n1, int n2, int n3, int S, int n_shells, const double *V) {
  int maxs = lattice == ORC_BCC ? ORC_BCC_MAX_SHELLS : lattice == ORC_FCC ? ORC_FCC_MAX_SHELLS
                                                                          : ORC_SC_MAX_SHELLS;
  if (lattice < 0 || lattice > 2 || n_shells < 1 || n_shells > maxs || S < 1 || S > 127) return NULL;
  orc_sys *s = (orc_sys *)calloc(1, sizeof(orc_sys));
  s->lattice = lattice; s->n1 = n1; s->n2 = n2; s->n3 = n3; s->S = S; s->n_shells = n_shells;
  s->gx = 2 * n1; s->gy = 2 * n2; s->gz = 2 * n3;
  if (lattice == ORC_SC) { s->wx = n1; s->wy = n2; s->wz = n3; }
  else { s->wx = s->gx; s->wy = s->gy; s->wz = s->gz; }
  s->n_atoms = (lattice == ORC_SC ? 8 : lattice == ORC_BCC ? 2 : 4) * n1 * n2 * n3;
  const int *st = lattice == ORC_BCC ? orc_bcc_start : lattice == ORC_FCC ? orc_fcc_start : orc_sc_start;
  const int *ct = lattice == ORC_BCC ? orc_bcc_count : lattice == ORC_FCC ? orc_fcc_count : orc_sc_count;
  s->off = lattice == ORC_BCC ? orc_bcc_off : lattice == ORC_FCC ? orc_fcc_off : orc_sc_off;
  for (int k = 0; k < n_shells; k++) { s->shell_start[k] = st[k]; s->shell_count[k] = ct[k]; s->z_total += ct[k]; }
  s->V = (double *)malloc(sizeof(double) * S * S * n_shells);
  memcpy(s->V, V, sizeof(double) * S * S * n_shells);
  return s;
}
void orc_sys_destroy(orc_sys *s) { if (s) { free(s->V); free(s); } }
int orc_sys_n_atoms(const orc_sys *s) { return s->n_atoms; }
int orc_sys_z_total(const orc_sys *s) { return s->z_total; }
long orc_sys_grid_size(const orc_sys *s) { return (long)s->gx * s->gy * s->gz; }

static inline int pmod(int a, int m) { int r = a % m; return r < 0 ? r + m : r; }
static inline long gidx(const orc_sys *s, int x, int y, int z) { return ((long)z * s->gy + y) * s->gx + x; }

/* Number of (dx,dy,dz) offsets and a copy of them, reference order, for n_shells shells */
int orc_sys_offsets(const orc_sys *s, int *out /* [z_total][4] = dx,dy,dz,shell */) {
  int n = 0;
  for (int k = 0; k < s->n_shells; k++)
    for (int j = 0; j < s->shell_count[k]; j++, n++) {
      const signed char *o = s->off[s->shell_start[k] + j];
      out[4 * n] = o[0]; out[4 * n + 1] = o[1]; out[4 * n + 2] = o[2]; out[4 * n + 3] = k;
    }
  return n;
}

/* ------------------------------------------------------------------------------------------
 * Hamiltonian
 * ---------------------------------------------------------------------------------------- */

/* <lattice>_shellK_energy with the centre species passed in: sequential f64 sum from 0.0 in
 * listed-neighbour order (e.g. src/bw_hamiltonian.f90:134-175).  species are 1-based in the grid. */
static double shell_energy(const orc_sys *s, const int8_t *g, int x, int y, int z, int k, int centre) {
  double e = 0.0;
  const signed char(*o)[3] = s->off + s->shell_start[k];
  const double *Vk = s->V + (long)k * s->S * s->S;
  for (int j = 0; j < s->shell_count[k]; j++) {
    int nb = g[gidx(s, pmod(x + o[j][0], s->wx), pmod(y + o[j][1], s->wy), pmod(z + o[j][2], s->wz))];
    e = e + Vk[(nb - 1) * s->S + (centre - 1)];
  }
  return e;
}

/* <lattice>_energy_Nshells == setup%nbr_energy: ((s1+s2)+s3)+... (src/bw_hamiltonian.f90:1003-1019) */
double orc_nbr_energy(const orc_sys *s, const int8_t *g, int x, int y, int z) {
  int centre = g[gidx(s, x, y, z)];
  double e = shell_energy(s, g, x, y, z, 0, centre);
  for (int k = 1; k < s->n_shells; k++) e = e + shell_energy(s, g, x, y, z, k, centre);
  return e;
}

/* pair_energy: nbr_energy(site1) + nbr_energy(site2), no 1/2 (src/bw_hamiltonian.f90:99-114) */
double orc_pair_energy(const orc_sys *s, const int8_t *g, const int *a, const int *b) {
  return orc_nbr_energy(s, g, a[0], a[1], a[2]) + orc_nbr_energy(s, g, b[0], b[1], b[2]);
}

/* total_energy: z-outer, y, x-inner sequential sum over occupied cells, then *0.5
 * (src/bw_hamiltonian.f90:58-81) */
double orc_total_energy(const orc_sys *s, const int8_t *g) {
  double e = 0.0;
  for (int z = 0; z < s->gz; z++)
    for (int y = 0; y < s->gy; y++)
      for (int x = 0; x < s->gx; x++) {
        if (g[gidx(s, x, y, z)] == 0) continue;
        e = e + orc_nbr_energy(s, g, x, y, z);
      }
  return 0.5 * e;
}

/* Per-site nbr_energy for every grid cell (0.0 on empty cells); test helper */
void orc_site_energies(const orc_sys *s, const int8_t *g, double *out) {
  for (int z = 0; z < s->gz; z++)
    for (int y = 0; y < s->gy; y++)
      for (int x = 0; x < s->gx; x++) {
        long i = gidx(s, x, y, z);
        out[i] = g[i] ? orc_nbr_energy(s, g, x, y, z) : 0.0;
      }
}

/* pair_swap (src/random_site.f90:238-250) */
static inline void pair_swap(const orc_sys *s, int8_t *g, const int *a, const int *b) {
  long ia = gidx(s, a[0], a[1], a[2]), ib = gidx(s, b[0], b[1], b[2]);
  int8_t t = g[ia]; g[ia] = g[ib]; g[ib] = t;
}

/* dE of swapping the occupants of a and b exactly as the Metropolis step computes it:
 * pair_energy(after) - pair_energy(before) (src/metropolis.F90:783-792).  grid is restored. */
double orc_pair_dE(const orc_sys *s, int8_t *g, const int *a, const int *b) {
  double e0 = orc_pair_energy(s, g, a, b);
  pair_swap(s, g, a, b);
  double e1 = orc_pair_energy(s, g, a, b);
  pair_swap(s, g, a, b);
  return e1 - e0;
}

/* Batched form: idx are flat grid indices; used by the GPU parity tests */
void orc_pair_dE_batch(const orc_sys *s, int8_t *g, long n, const int32_t *i1, const int32_t *i2, double *dE) {
  for (long t = 0; t < n; t++) {
    int a[3] = {i1[t] % s->gx, (i1[t] / s->gx) % s->gy, i1[t] / (s->gx * s->gy)};
    int b[3] = {i2[t] % s->gx, (i2[t] / s->gx) % s->gy, i2[t] / (s->gx * s->gy)};
    dE[t] = orc_pair_dE(s, g, a, b);
  }
}

/* ------------------------------------------------------------------------------------------
 * Site proposal (src/random_site.f90).  RNG order: z, then x, then y.
 * ---------------------------------------------------------------------------------------- */
static const signed char sc_nbrs[6][3] = {{0,0,1},{0,1,0},{1,0,0},{0,0,-1},{0,-1,0},{-1,0,0}};           /* :29-35 */
static const signed char bcc_nbrs[8][3] = {{1,1,1},{1,1,-1},{1,-1,1},{1,-1,-1},{-1,1,1},{-1,1,-1},{-1,-1,1},{-1,-1,-1}}; /* :38-46 */
static const signed char fcc_nbrs[12][3] = {{0,1,1},{0,1,-1},{0,-1,1},{0,-1,-1},{1,1,0},{1,-1,0},{1,0,1},{1,0,-1},
                                            {-1,1,0},{-1,-1,0},{-1,0,1},{-1,0,-1}};                       /* :49-61 */

void orc_random_site(const orc_sys *s, orc_mt *g, int *site) {
  int x1, y1, z1; /* 1-based, as the reference computes them */
  if (s->lattice == ORC_SC) {            /* :74-84 (x, y, z order for simple cubic) */
    x1 = (int)floor(orc_mt_genrand(g) * 2.0 * (double)s->n1) + 1;
    y1 = (int)floor(orc_mt_genrand(g) * 2.0 * (double)s->n2) + 1;
    z1 = (int)floor(orc_mt_genrand(g) * 2.0 * (double)s->n3) + 1;
  } else if (s->lattice == ORC_BCC) {    /* :127-139 */
    z1 = (int)floor(2.0 * orc_mt_genrand(g) * (double)s->n3) + 1;
    x1 = 2 * (int)floor(orc_mt_genrand(g) * (double)s->n1) + 2 - (z1 % 2);
    y1 = 2 * (int)floor(orc_mt_genrand(g) * (double)s->n2) + 2 - (z1 % 2);
  } else {                               /* :182-193 */
    z1 = (int)floor(2.0 * orc_mt_genrand(g) * (double)s->n3) + 1;
    x1 = (int)floor(2.0 * orc_mt_genrand(g) * (double)s->n1) + 1;
    y1 = 2 * (int)floor(orc_mt_genrand(g) * (double)s->n2) + 1 + pmod(x1 - (z1 % 2), 2);
  }
  site[0] = x1 - 1; site[1] = y1 - 1; site[2] = z1 - 1;
}

void orc_random_nbr(const orc_sys *s, orc_mt *g, const int *site, int *nbr) { /* :97-116,152-171,206-225 */
  int z1n = s->lattice == ORC_SC ? 6 : s->lattice == ORC_BCC ? 8 : 12;
  const signed char(*t)[3] = s->lattice == ORC_SC ? sc_nbrs : s->lattice == ORC_BCC ? bcc_nbrs : fcc_nbrs;
  int n = (int)floor((double)z1n * orc_mt_genrand(g));
  nbr[0] = pmod(site[0] + t[n][0], s->gx);
  nbr[1] = pmod(site[1] + t[n][1], s->gy);
  nbr[2] = pmod(site[2] + t[n][2], s->gz);
}

/* ------------------------------------------------------------------------------------------
 * Initial configuration (src/initialise.F90:434-617)
 * ---------------------------------------------------------------------------------------- */

/* Quotas + thresholds as HEAD derives them (:468-506).  conc_io[0..S] with conc_io[0] = 0.0
 * (species_concentrations(0:), src/io.f90:242-246); numbers may be NULL / not summing to N. */
int orc_species_quotas(const orc_sys *s, double *conc_io, const int64_t *numbers, int64_t *count) {
  int S = s->S; int64_t n_sites = s->n_atoms, sum = 0;
  if (numbers) for (int i = 0; i < S; i++) sum += numbers[i];
  if (numbers && sum == s->n_atoms) {
    for (int i = 0; i < S; i++) {
      count[i] = numbers[i];
      conc_io[i + 1] = (double)((float)count[i] / (float)s->n_atoms); /* single precision, :471 */
    }
  } else {
    int64_t tot = 0;
    for (int i = 0; i < S; i++) {
      /* nint(real(n_sites)*conc): real(n_sites) is single precision, product is double (:482) */
      count[i] = (int64_t)llround((double)(float)n_sites * conc_io[i + 1]);
      tot += count[i];
    }
    int64_t round_err = tot - n_sites, inc = round_err > 0 ? -1 : 1, chk = round_err;
    while (chk != 0)
      for (int j = 0; j < S; j++) {
        if (count[j] == 0) continue;
        if (chk == 0) continue;
        count[j] += inc; chk += inc;
      }
  }
  sum = 0;
  for (int i = 0; i < S; i++) sum += count[i];
  return sum == n_sites ? 0 : -1;
}

/* Draw until a species whose threshold interval contains r still has quota (:528-545 etc.).
 * No early exit from the species loop, exactly like the reference. */
static void fill_site(const orc_sys *s, int8_t *cell, const double *cum, const int64_t *count,
                      int64_t *check, orc_mt *g) {
  while (*cell == 0) {
    double r = orc_mt_genrand(g);
    for (int l = 1; l <= s->S; l++)
      if (r >= cum[l - 1] && r <= cum[l])
        if (check[l - 1] < count[l - 1]) { *cell = (int8_t)l; check[l - 1]++; }
  }
}

/* conc[0..S] (conc[0]=0), count[S] quotas.  Cells visited z-outer / y / x-inner over occupied
 * positions.  Returns 0, or -1 on bad quotas. */
int orc_initial_setup(const orc_sys *s, const double *conc, const int64_t *count, orc_mt *g, int8_t *grid) {
  int S = s->S;
  double cum[130]; int64_t check[128];
  int64_t tot = 0;
  for (int i = 0; i < S; i++) { tot += count[i]; check[i] = 0; }
  if (tot != s->n_atoms) return -1;
  /* sum(species_concentrations(0:l)) : sequential sum starting from element 0 */
  for (int l = 0; l <= S; l++) { double c = 0.0; for (int i = 0; i <= l; i++) c += conc[i]; cum[l] = c; }
  memset(grid, 0, (size_t)orc_sys_grid_size(s));
  if (s->lattice == ORC_SC) {                       /* :525-545 */
    for (int k = 1; k <= s->gz; k++) for (int j = 1; j <= s->gy; j++) for (int i = 1; i <= s->gx; i++)
      fill_site(s, &grid[gidx(s, i - 1, j - 1, k - 1)], cum, count, check, g);
  } else if (s->lattice == ORC_BCC) {               /* :547-572 */
    for (int k = 1; k <= s->gz; k++) for (int j = 1; j <= s->gy / 2; j++) for (int i = 1; i <= s->gx / 2; i++) {
      int x1 = 2 * i - (k % 2), y1 = 2 * j - (k % 2);
      fill_site(s, &grid[gidx(s, x1 - 1, y1 - 1, k - 1)], cum, count, check, g);
    }
  } else {                                          /* :574-610 */
    for (int k = 1; k <= s->gz; k++) for (int j = 1; j <= s->gy; j++) for (int i = 1; i <= s->gx / 2; i++) {
      int x1 = 2 * i - (k % 2) * (j % 2) - ((k + 1) % 2) * ((j + 1) % 2);
      fill_site(s, &grid[gidx(s, x1 - 1, j - 1, k - 1)], cum, count, check, g);
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * Metropolis (src/metropolis.F90:751-891)
 * ---------------------------------------------------------------------------------------- */
int orc_mc_step(const orc_sys *s, int8_t *grid, orc_mt *g, double beta, int nbr_swap) {
  int a[3], b[3];
  orc_random_site(s, g, a);
  if (nbr_swap) orc_random_nbr(s, g, a, b); else orc_random_site(s, g, b);
  int8_t s1 = grid[gidx(s, a[0], a[1], a[2])], s2 = grid[gidx(s, b[0], b[1], b[2])];
  if (s1 == s2) return 1;                                   /* :774-777 */
  double e_unswapped = orc_pair_energy(s, grid, a, b);      /* :783 */
  pair_swap(s, grid, a, b);                                 /* :786 */
  double e_swapped = orc_pair_energy(s, grid, a, b);        /* :789 */
  double delta_e = e_swapped - e_unswapped;                 /* :792 */
  if (delta_e < 0.0) return 1;                              /* :796 */
  if (orc_mt_genrand(g) < exp(-beta * delta_e)) return 1;   /* :802 */
  pair_swap(s, grid, a, b);                                 /* :810 */
  return 0;
}

/* n_trials calls of mc_step (the k-loop, src/metropolis.F90:350-354); returns #accepted */
int64_t orc_metropolis_trials(const orc_sys *s, int8_t *grid, orc_mt *g, double beta, int64_t n_trials, int nbr_swap) {
  int64_t acc = 0;
  for (int64_t t = 0; t < n_trials; t++) acc += orc_mc_step(s, grid, g, beta, nbr_swap);
  return acc;
}

/* One temperature of metropolis_simulated_annealing's sampling loop (src/metropolis.F90:340-437)
 * with energies sampled every n_sample_steps.  energies_out (may be NULL) receives the
 * full energy after each sweep.  out[0]=<E>/atom, out[1]=C, out[2]=acceptance. */
void orc_metropolis_sample(const orc_sys *s, int8_t *grid, orc_mt *g, double temp, double k_b_in_ry,
                           int64_t n_mc_steps, int64_t n_sample_steps, int nbr_swap,
                           double *energies_out, double *out) {
  double sim_temp = temp * k_b_in_ry, beta = 1.0 / sim_temp;       /* :204-206 */
  int64_t n_sweeps = n_mc_steps / n_sample_steps, n_sweep_steps = n_mc_steps / n_sweeps; /* :343-344 */
  /* n_save_energy = floor(real(n_mc_steps)/real(n_sample_steps)) in single precision (:155-156) */
  int n_save_energy = (int)floorf((float)n_mc_steps / (float)n_sample_steps);
  double step_E = 0.0, step_Esq = 0.0, acceptance = 0.0;
  for (int64_t i = 0; i < n_sweeps; i++) {
    for (int64_t k = 0; k < n_sweep_steps; k++) acceptance = acceptance + orc_mc_step(s, grid, g, beta, nbr_swap);
    double e = orc_total_energy(s, grid);
    step_E = step_E + e; step_Esq = step_Esq + e * e;
    if (energies_out) energies_out[i] = e;
  }
  out[2] = acceptance / (double)(float)n_mc_steps;                     /* :412 real() is single */
  out[0] = step_E / n_save_energy / s->n_atoms;                        /* :416 */
  double C = (step_Esq / n_save_energy - (step_E / n_save_energy) * (step_E / n_save_energy)) / (sim_temp * temp) / s->n_atoms;
  if (C < 0.0) C = 0.0;
  if (temp <= 0.0) C = 0.0;
  out[1] = C;
}

/* ------------------------------------------------------------------------------------------
 * Wang-Landau sweeps for ONE walker (src/wang-landau.F90:515-626; the allreduce at :628-631
 * is the caller's business).  lng[bins], hist[win_hi-win_lo+1]; win_lo/win_hi are the 1-based
 * mpi_start_idx / mpi_end_idx; bin_edges[bins+1].  radial_densities sampling (:574-592) and
 * config_store (:607-610) are not part of the arithmetic and are omitted.
 * ---------------------------------------------------------------------------------------- */
static int bin_index(double e, const double *edges, int bins) { /* :515-523 */
  double range = edges[bins] - edges[0];
  return (int)(((e - edges[0]) / range) * (double)bins) + 1;
}
int orc_bin_index(double e, const double *edges, int bins) { return bin_index(e, edges, bins); }

int64_t orc_wl_sweeps(const orc_sys *s, int8_t *grid, orc_mt *g, double *lng, double *hist,
                      const double *edges, int bins, int win_lo, int win_hi, double wl_f,
                      int64_t n_trials, int nbr_swap, double *e_final) {
  double e_unswapped = orc_total_energy(s, grid), e_swapped = e_unswapped; /* :547-548 */
  int64_t accepted = 0;
  int hist_every = (int)(0.02 * (double)(float)s->n_atoms);            /* INT(0.02_real64*REAL(n_atoms)) :605 */
  for (int64_t i = 1; i <= n_trials; i++) {
    int a[3], b[3];
    orc_random_site(s, g, a);
    if (nbr_swap) orc_random_nbr(s, g, a, b); else orc_random_site(s, g, b);
    int8_t s1 = grid[gidx(s, a[0], a[1], a[2])], s2 = grid[gidx(s, b[0], b[1], b[2])];
    e_swapped = e_unswapped;
    double pair_unswapped = orc_pair_energy(s, grid, a, b), pair_swapped = pair_unswapped; /* :561-562 */
    pair_swap(s, grid, a, b);
    if (s1 != s2) {
      pair_swapped = orc_pair_energy(s, grid, a, b);
      e_swapped = e_unswapped - pair_unswapped + pair_swapped;        /* :568 */
    }
    int ibin = bin_index(e_unswapped, edges, bins), jbin = bin_index(e_swapped, edges, bins);
    if (jbin > win_lo - 1 && jbin < win_hi + 1) {                      /* :595 */
      if (log(orc_mt_genrand(g)) < (lng[ibin - 1] - lng[jbin - 1])) { /* :598 */
        accepted++; e_unswapped = e_swapped;
      } else { pair_swap(s, grid, a, b); jbin = ibin; }
    } else { jbin = ibin; pair_swap(s, grid, a, b); }                  /* :614-617 */
    if (hist_every > 0 && i % hist_every == 0) hist[jbin - win_lo] += 1.0; /* :605-606 / :618-619 */
    lng[jbin - 1] += wl_f;                                             /* :612 / :624 */
  }
  if (e_final) *e_final = e_unswapped;
  return accepted;
}

/* ------------------------------------------------------------------------------------------
 * Nested sampling (src/nested_sampling.f90:78-200).  walkers: K grids back to back.
 * culled[n_iter] receives ener_limit per iteration (the numbers written to unit 35, :121).
 * conc/count as for orc_initial_setup.
 * ---------------------------------------------------------------------------------------- */
int orc_nested_sampling(const orc_sys *s, orc_mt *g, const double *conc, const int64_t *count,
                        int K, int n_steps, int n_iter, int8_t *walkers, double *energies, double *culled) {
  long G = orc_sys_grid_size(s);
  for (int w = 0; w < K; w++) {                                        /* :78-97 */
    if (orc_initial_setup(s, conc, count, g, walkers + w * G)) return -1;
    double rnde = orc_mt_genrand(g);
    energies[w] = orc_total_energy(s, walkers + w * G) + rnde * (double)1e-8f; /* default-real literal :95 */
  }
  int n_at = s->n1 * s->n2 * s->n3 * 1 * s->S;                         /* :84 (sic) */
  int extra_steps = 0, n_acc = 0;
  for (int it = 1; it <= n_iter; it++) {                               /* :114 */
    int i_max = 0;
    for (int w = 1; w < K; w++) if (energies[w] > energies[i_max]) i_max = w; /* maxloc: first max */
    double ener_limit = energies[i_max];
    culled[it - 1] = ener_limit;
    if (it % (int)(K / 2.0) == 0) {                                    /* :129-144 */
      if (((float)n_acc < (float)n_at * 0.05f) && (extra_steps < n_steps * 100)) extra_steps += n_steps;
    }
    double rnd = orc_mt_genrand(g);                                    /* :149 */
    int irnd = (int)ceil(rnd * K);                                     /* 1-based */
    if (irnd < 1) irnd = 1; /* rnd == 0.0 would index walker 0 in the reference (out of bounds) */
    if (irnd - 1 != i_max) memcpy(walkers + i_max * G, walkers + (irnd - 1) * G, (size_t)G);
    energies[i_max] = energies[irnd - 1];
    int8_t *wk = walkers + i_max * G;
    n_acc = 0;
    for (int st = 1; st <= n_steps + extra_steps; st++) {              /* :157-192 */
      int a[3], b[3];
      orc_random_site(s, g, a);
      int8_t s1 = wk[gidx(s, a[0], a[1], a[2])], s2;
      do { orc_random_site(s, g, b); s2 = wk[gidx(s, b[0], b[1], b[2])]; } while (s1 == s2);
      double e_unswapped = orc_pair_energy(s, wk, a, b);
      pair_swap(s, wk, a, b);
      double e_swapped = orc_pair_energy(s, wk, a, b);
      double delta_e = e_swapped - e_unswapped;
      if (energies[i_max] + delta_e < ener_limit) { energies[i_max] = energies[i_max] + delta_e; n_acc++; }
      else pair_swap(s, wk, a, b);
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * Observables (src/analytics.f90)
 * ---------------------------------------------------------------------------------------- */

/* lattice_shells (:205-275): sorted distinct distances sqrt(x^2+y^2+z^2) (single-precision
 * sqrt of an integer-valued real) from the origin cell to every occupied cell. */
static int cmp_d(const void *a, const void *b) { double x = *(const double *)a, y = *(const double *)b; return (x > y) - (x < y); }
void orc_lattice_shells(const orc_sys *s, const int8_t *grid, int wc_range, double *shells) {
  long G = orc_sys_grid_size(s), l = 0;
  double *all = (double *)calloc((size_t)G + 1, sizeof(double)); /* all_shells(8*n1*n2*n3+1) = 0 */
  for (int k = 0; k < s->gz; k++) for (int j = 0; j < s->gy; j++) for (int i = 0; i < s->gx; i++) {
    if (grid[gidx(s, i, j, k)] == 0) continue;
    all[l++] = (double)sqrtf((float)(k * k) + (float)(j * j) + (float)(i * i));
  }
  qsort(all, (size_t)G + 1, sizeof(double), cmp_d);
  for (int i = 0; i < wc_range; i++) shells[i] = 0.0;
  int n = 0;
  for (long i = 0; i < G && n < wc_range; i++) {
    if (fabs(all[i] - all[i + 1]) < 1e-3) continue;
    shells[n++] = all[i];
  }
  free(all);
}

/* radial_densities (:293-404): cube scan of half-width min(n,5), distances in f64 vs single-precision shell radii,
 * rho(i,j,l) = #{j-type on shell l around i-type} / N_i.  out[(l*S + j)*S + i] (Fortran order). */
void orc_radial_densities(const orc_sys *s, const int8_t *grid, int wc_range, const double *shells, double *out) {
  int S = s->S; long cnt[128] = {0};
  memset(out, 0, sizeof(double) * S * S * wc_range);
  long G = orc_sys_grid_size(s);
  for (long i = 0; i < G; i++) if (grid[i] > 0) cnt[grid[i] - 1]++;
  int l1 = s->n1 < 5 ? s->n1 : 5, l2 = s->n2 < 5 ? s->n2 : 5, l3 = s->n3 < 5 ? s->n3 : 5;
  for (int i3 = 0; i3 < s->gz; i3++) for (int i2 = 0; i2 < s->gy; i2++) for (int i1 = 0; i1 < s->gx; i1++) {
    int si = grid[gidx(s, i1, i2, i3)];
    if (!si) continue;
    for (int jj3 = i3 - l3; jj3 <= i3 + l3; jj3++) { int j3 = pmod(jj3, s->gz);
      for (int jj2 = i2 - l2; jj2 <= i2 + l2; jj2++) { int j2 = pmod(jj2, s->gy);
        for (int jj1 = i1 - l1; jj1 <= i1 + l1; jj1++) { int j1 = pmod(jj1, s->gx);
          int sj = grid[gidx(s, j1, j2, j3)];
          if (!sj) continue;
          /* d_x etc. are real64 in the reference; nint() rounds half away from zero == round() */
          double dx = (double)(i1 - j1), dy = (double)(i2 - j2), dz = (double)(i3 - j3);
          dx = dx - (double)s->gx * round(dx / (double)s->gx);
          dy = dy - (double)s->gy * round(dy / (double)s->gy);
          dz = dz - (double)s->gz * round(dz / (double)s->gz);
          double distance = sqrt(dx * dx + dy * dy + dz * dz);
          for (int l = 0; l < wc_range; l++)
            if (fabs(distance - shells[l]) < 1e-3) out[((long)l * S + (sj - 1)) * S + (si - 1)] += 1.0;
        } } }
  }
  for (int l = 0; l < wc_range; l++) for (int j = 0; j < S; j++) for (int i = 0; i < S; i++)
    out[((long)l * S + j) * S + i] /= (double)cnt[i];   /* r_densities(j,:,i)/particle_counts(j) :398-402 */
}

/* Integer pair counts over the tabulated coordination shells (test helper for the GPU SRO
 * kernel): cnt[(l*S + j)*S + i] = # of j-type neighbours on coordination shell l (l=0: the
 * site itself) of i-type atoms; species_count[S]. */
void orc_radial_counts(const orc_sys *s, const int8_t *grid, int wc_range, int64_t *cnt, int64_t *species_count) {
  int S = s->S;
  memset(cnt, 0, sizeof(int64_t) * S * S * wc_range);
  memset(species_count, 0, sizeof(int64_t) * S);
  for (int z = 0; z < s->gz; z++) for (int y = 0; y < s->gy; y++) for (int x = 0; x < s->gx; x++) {
    int si = grid[gidx(s, x, y, z)];
    if (!si) continue;
    species_count[si - 1]++;
    cnt[((long)0 * S + (si - 1)) * S + (si - 1)]++;
    for (int l = 1; l < wc_range; l++) {
      int k = l - 1;
      const signed char(*o)[3] = s->off + (s->lattice == ORC_BCC ? orc_bcc_start[k] : s->lattice == ORC_FCC ? orc_fcc_start[k] : orc_sc_start[k]);
      int n = s->lattice == ORC_BCC ? orc_bcc_count[k] : s->lattice == ORC_FCC ? orc_fcc_count[k] : orc_sc_count[k];
      for (int j = 0; j < n; j++) {
        int sj = grid[gidx(s, pmod(x + o[j][0], s->gx), pmod(y + o[j][1], s->gy), pmod(z + o[j][2], s->gz))];
        if (sj) cnt[((long)l * S + (sj - 1)) * S + (si - 1)]++;
      }
    }
  }
}

/* ---- Wang-Landau control plane (host arithmetic; parity UNPINNED: the reference ships no golden for it) ---- */

/* compute_mean_energy, src/wang-landau.F90:457-477. out[300][2] = (<E>, beta). */
void orc_wl_mean_energy(const double *lng, const double *edges, int bins, double bin_width, double k_b_in_ry,
                        double *out) {
  double mx = lng[0];
  for (int b = 1; b < bins; b++) if (lng[b] > mx) mx = lng[b];
  double *prob = (double *)malloc(sizeof(double) * (size_t)bins);
  for (int itemp = 1; itemp <= 300; itemp++) {
    const double beta = 1.0 / (k_b_in_ry * itemp * 10.0);
    double pm = 0.0;
    for (int b = 0; b < bins; b++) {
      prob[b] = (lng[b] - mx) - beta * (edges[b] + 0.5 * bin_width);
      if (b == 0 || prob[b] > pm) pm = prob[b];
    }
    double s = 0.0;
    for (int b = 0; b < bins; b++) { prob[b] = exp(prob[b] - pm); s += prob[b]; }
    double e = 0.0;
    for (int b = 0; b < bins; b++) e += (edges[b] + 0.5 * bin_width) * (prob[b] / s);
    out[2 * (itemp - 1)] = e;
    out[2 * (itemp - 1) + 1] = beta;
  }
  free(prob);
}

/* mpi_window_optimise (rank-0 arithmetic), src/wang-landau.F90:1224-1311; sort_descending :1686-1701.
   iv[W][2] in/out (1-based inclusive), prev[W] in/out (diffusion_prev). Returns 0, or 1 if W < 2 (nothing done). */
int orc_wl_window_optimise(int iter, int W, int64_t *iv, const double *mc_steps, double *prev, int bins) {
  if (W < 2) return 1;
  double alpha = 0.8 * pow(0.8, (double)(iter - 1));
  if (iter == 0) alpha = 1.0;
  const double w_min = 0.02;
  double *wmc = (double *)malloc(sizeof(double) * (size_t)W), *frac = (double *)malloc(sizeof(double) * (size_t)W);
  int *nb = (int *)malloc(sizeof(int) * (size_t)W), *idx = (int *)malloc(sizeof(int) * (size_t)W);
  double s = 0.0;
  for (int i = 0; i < W; i++) {
    const long first = (long)iv[2 * i], last = (long)iv[2 * i + 1];
    wmc[i] = 1.0 / (mc_steps[i] / (double)(float)labs(first - last + 1));
    s += wmc[i];
  }
  double sf = 0.0;
  for (int i = 0; i < W; i++) { frac[i] = alpha * (wmc[i] / s) + (1.0 - alpha) * prev[i]; sf += frac[i]; }
  for (int i = 0; i < W; i++) { frac[i] /= sf; prev[i] = frac[i]; }
  double tot = 0.0, sum_free = 0.0;
  int any_free = 0;
  for (int i = 0; i < W; i++) { if (frac[i] < w_min) frac[i] = w_min; tot += frac[i]; }
  if (fabs(tot - 1.0) > 1.0e-12) {
    for (int i = 0; i < W; i++) if (frac[i] > w_min) { any_free = 1; sum_free += frac[i]; }
    if (any_free && sum_free > 0.0) {
      const double scale = (1.0 - tot) / sum_free;
      for (int i = 0; i < W; i++) if (frac[i] > w_min) frac[i] = frac[i] + frac[i] * scale;
    }
  }
  tot = 0.0;
  for (int i = 0; i < W; i++) tot += frac[i];
  int min_bins = (int)(w_min * bins);
  if (min_bins < 2) min_bins = 2;
  long have = 0;
  for (int i = 0; i < W; i++) {
    nb[i] = (int)lround((double)(float)bins * (frac[i] / tot));      /* NINT: half away from zero */
    if (nb[i] < min_bins) nb[i] = min_bins;
    have += nb[i];
  }
  if (have != bins) {
    for (int i = 0; i < W; i++) idx[i] = i;
    for (int i = 0; i < W - 1; i++)
      for (int j = i + 1; j < W; j++)
        if (nb[idx[i]] < nb[idx[j]]) { int t = idx[i]; idx[i] = idx[j]; idx[j] = t; }
    long diff = bins - have;
    for (long i = 0; diff != 0 && i < 8L * W * bins; i++) {
      const int j = idx[i % W];
      if (diff > 0) { nb[j]++; diff--; }
      else if (nb[j] > min_bins) { nb[j]--; diff++; }
    }
  }
  iv[1] = nb[0];
  for (int i = 1; i < W; i++) { iv[2 * i] = iv[2 * (i - 1) + 1] + 1; iv[2 * i + 1] = iv[2 * i] + nb[i] - 1; }
  iv[2 * (W - 1)] = iv[2 * (W - 2) + 1] + 1;
  iv[2 * (W - 1) + 1] = bins;
  free(wmc); free(frac); free(nb); free(idx);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * Wang-Landau: window entry, stitching, replica exchange (src/wang-landau.F90:643-741, 1147-1194,
 * 1392-1519).  Parity: the reference ships no golden for these (its own run is not reproducible:
 * MPI_ANY_SOURCE timing decides who enters a window first), so they are pinned by construction --
 * line-by-line restatements, compared with the GPU replay kernel / the host implementations.
 * ---------------------------------------------------------------------------------------- */

/* enter_energy_window (:643-741) for ONE walker (no other rank of the window ever reports in, i.e.
 * `flag` stays false).  min_e / max_e = MINVAL / MAXVAL(mpi_bin_edges) in Ry per cell; energy_min /
 * energy_max are the single-precision inputs of wl_input.inp (meV/atom).  One loop iteration = one
 * pass of the `do while(.True.)` body; the walk stops when it leaves through :685-703 (returns 1) or
 * after max_iters iterations (returns 0).  Reference behaviours kept: i_steps restarts at 0 and the
 * lattice is re-randomised with initial_setup every n_atoms*250 iterations WITHOUT refreshing the
 * running energy (:669-674); the window test uses the running energy first and only then the exact
 * one (:685-690), an iteration that fails the second test makes no trial (`cycle`).
 * Outputs: *e_final = e_unswapped on exit, *n_iters = loop iterations executed. */
int orc_wl_enter_energy_window(const orc_sys *s, int8_t *grid, orc_mt *g, const double *conc,
                               const int64_t *count, double min_e, double max_e, float energy_min,
                               float energy_max, int64_t max_iters, double *e_final, int64_t *n_iters) {
  const double target_energy = (min_e + max_e) / 2.0;                       /* :663 */
  const double condition = fabs(max_e - min_e) * 0.1;                       /* :664 */
  /* 0.0025*ABS(energy_max - energy_min)*n_atoms: default-real (single) arithmetic, then / (Ry_to_eV*1000) in double (:726-727) */
  float sf = 0.0025f * fabsf(energy_max - energy_min);
  sf = sf * (float)s->n_atoms;
  const double sigma = (double)sf / (13.605693122 * 1000.0);
  const double denom = 2.0 * (sigma * sigma);
  double e_unswapped = orc_total_energy(s, grid);                           /* :669 */
  int64_t i_steps = 0, it = 0;
  const int64_t period = (int64_t)s->n_atoms * 50 * 5;
  int entered = 0;
  while (it < max_iters) {
    it++;
    i_steps++;
    if (i_steps % period == 0) {                                            /* :677-680 */
      i_steps = 0;
      orc_initial_setup(s, conc, count, g, grid);
    }
    if (e_unswapped < max_e - condition && e_unswapped > min_e + condition) {   /* :691 */
      e_unswapped = orc_total_energy(s, grid);
      if (e_unswapped < max_e - condition && e_unswapped > min_e + condition) { entered = 1; break; }
      continue;                                                             /* cycle */
    }
    int a[3], b[3];
    orc_random_site(s, g, a);                                               /* :710-711 */
    orc_random_site(s, g, b);
    int8_t s1 = grid[gidx(s, a[0], a[1], a[2])], s2 = grid[gidx(s, b[0], b[1], b[2])];
    if (s1 != s2) {                                                         /* :718 */
      double pair_unswapped = orc_pair_energy(s, grid, a, b);
      pair_swap(s, grid, a, b);
      double pair_swapped = orc_pair_energy(s, grid, a, b);
      double e_swapped = e_unswapped - pair_unswapped + pair_swapped;       /* :724 */
      double d1 = e_swapped - target_energy, d0 = e_unswapped - target_energy;
      double delta_e = (d1 * d1 - d0 * d0) / denom;                         /* :727-729 */
      if (log(orc_mt_genrand(g)) < -delta_e) e_unswapped = e_swapped;       /* :731 */
      else pair_swap(s, grid, a, b);
    }
  }
  if (e_final) *e_final = e_unswapped;
  if (n_iters) *n_iters = it;
  return entered;
}

/* dos_combine (:1147-1194) as rank 0 evaluates it.  lng[W][bins]: the window-averaged ln g of every
 * window (what window_rank_index(i,1) sends); win[W][2]: window_indices (1-based, inclusive).
 * out[bins].  Kept: beta_index is NOT reset between windows (an empty overlap loop re-uses the last
 * one); the stitch loop starts AT beta_index, so comb(beta_index) is first rewritten to
 * (buf + comb) - buf and the rewritten value is what the later bins see. */
void orc_wl_dos_combine(const double *lng, const int64_t *win, int W, int bins, double *out) {
  int beta_index = 0;
  memcpy(out, lng, sizeof(double) * (size_t)bins);                          /* :1159 */
  for (int i = 2; i <= W; i++) {
    const double *buf = lng + (size_t)(i - 1) * bins;
    const int start = (int)win[2 * (i - 1)], end = (int)win[2 * (i - 1) + 1];
    double beta_diff = 1.7976931348623157e308;                              /* HUGE(1.0_real64) */
    const int jmax = (int)(win[2 * (i - 2) + 1] - win[2 * (i - 1)] - 1);
    for (int j = 0; j <= jmax; j++) {                                       /* :1174-1181 */
      double beta_original = out[start + j + 1 - 1] - out[start + j - 1];
      double beta_merge = buf[start + j + 1 - 1] - buf[start + j - 1];
      if (fabs(beta_original - beta_merge) < beta_diff) {
        beta_diff = fabs(beta_original - beta_merge);
        beta_index = start + j + 1;
      }
    }
    for (int j = beta_index; j <= end; j++)                                 /* :1183-1185 */
      out[j - 1] = buf[j - 1] + out[beta_index - 1] - buf[beta_index - 1];
  }
  double mn = out[0];
  for (int b = 1; b < bins; b++) if (out[b] < mn) mn = out[b];
  for (int b = 0; b < bins; b++) out[b] = out[b] - mn;                      /* :1190 */
}

/* shuffle_rows (:1512-1519), rank 0's MT stream: rows = walker ranks with their overlap location */
static void shuffle_rows(int *rows /* [n][2] */, int n, orc_mt *g) {
  for (int i = n; i >= 2; i--) {
    int r = 1 + (int)(orc_mt_genrand(g) * (double)i);
    if (r > n) r = n;
    int t0 = rows[2 * (i - 1)], t1 = rows[2 * (i - 1) + 1];
    rows[2 * (i - 1)] = rows[2 * (r - 1)]; rows[2 * (i - 1) + 1] = rows[2 * (r - 1) + 1];
    rows[2 * (r - 1)] = t0; rows[2 * (r - 1) + 1] = t1;
  }
}

/* replica_exchange (:1392-1501) for P = W*num_walkers ranks (rank r = walker r%num_walkers of window
 * r/num_walkers + 1), given every rank's total energy.  lng[P][bins] is each rank's wl_logdos (window-averaged in the
 * caller), win[W][2] the window_indices, mts[P] the ranks' MT streams (rank 0 shuffles, :1441-1442; the LOWER walker of
 * a matched pair draws the acceptance uniform, :1482).  Bins are computed once from the energies at entry and not
 * refreshed after an exchange (:1399, 1405-1431); `accept` is never reset to .False. once it is .True. on a rank
 * (declared at :1404, set :1483), so a later rejected pair on the same rank still reports the stale .True. to its partner
 * -- with one overlap region per walker this cannot happen (a walker is matched at most once per call), kept anyway.
 * Output: pairs[n][2] = (lower rank, upper rank) of the exchanges carried out, in order; returns n (<= P). */
int orc_wl_replica_exchange(const double *energies, const double *lng, const int64_t *win, int W, int num_walkers,
                            const double *edges, int bins, orc_mt *mts, int *pairs) {
  const int P = W * num_walkers;
  int *ibin = (int *)malloc(sizeof(int) * (size_t)P), *loc = (int *)malloc(sizeof(int) * (size_t)P);
  int *lower = (int *)malloc(sizeof(int) * 2 * (size_t)num_walkers), *upper = (int *)malloc(sizeof(int) * 2 * (size_t)num_walkers);
  int *exch = (int *)malloc(sizeof(int) * 2 * (size_t)num_walkers);
  char *accept = (char *)calloc((size_t)P, 1);
  for (int r = 0; r < P; r++) {
    const int q = r / num_walkers + 1;                                      /* mpi_index */
    const int ib = bin_index(energies[r], edges, bins);
    ibin[r] = ib;
    int lo = 0, up = 0;
    if (q > 1) lo = (ib < win[2 * (q - 2) + 1] + 1) && (ib > win[2 * (q - 1)] - 1);          /* :1409-1413 */
    if (q < W) up = (ib > win[2 * q] - 1) && (ib < win[2 * (q - 1) + 1] + 1);                /* :1415-1419 */
    loc[r] = up ? q : (lo ? q - 1 : 0);                                     /* :1421-1427 */
  }
  int n = 0;
  for (int i = 1; i <= W - 1; i++) {
    for (int k = 0; k < num_walkers; k++) {
      lower[2 * k] = (i - 1) * num_walkers + k; lower[2 * k + 1] = loc[(i - 1) * num_walkers + k];
      upper[2 * k] = i * num_walkers + k; upper[2 * k + 1] = loc[i * num_walkers + k];
      exch[2 * k] = exch[2 * k + 1] = -1;
    }
    shuffle_rows(lower, num_walkers, &mts[0]);
    shuffle_rows(upper, num_walkers, &mts[0]);
    int ei = 0;
    for (int j = 0; j < num_walkers; j++) {                                 /* :1444-1462 */
      if (lower[2 * j + 1] == 0) continue;
      for (int k = 0; k < num_walkers; k++) {
        if (upper[2 * k + 1] == 0) continue;
        /* after a match lower(j,:) = 0, so the comparison below fails for the remaining k (upper rows with 0 are skipped) */
        if (lower[2 * j + 1] == upper[2 * k + 1]) {
          exch[2 * ei] = lower[2 * j]; exch[2 * ei + 1] = upper[2 * k]; ei++;
          lower[2 * j] = lower[2 * j + 1] = 0; upper[2 * k] = upper[2 * k + 1] = 0;
        }
      }
    }
    int cnt = 0;
    for (int j = 0; j < num_walkers; j++) if (exch[2 * j] > -1) cnt++;      /* COUNT(overlap_exchange(:,1) > -1) */
    for (int j = 0; j < cnt; j++) {                                         /* :1470-1497 */
      const int a = exch[2 * j], b = exch[2 * j + 1];
      const int jb = bin_index(energies[b], edges, bins);
      const double *la = lng + (size_t)a * bins;
      if (orc_mt_genrand(&mts[a]) < exp(la[ibin[a] - 1] - la[jb - 1])) accept[a] = 1;        /* :1482-1483 */
      if (accept[a]) { pairs[2 * n] = a; pairs[2 * n + 1] = b; n++; }
    }
  }
  free(ibin); free(loc); free(lower); free(upper); free(exch); free(accept);
  return n;
}
