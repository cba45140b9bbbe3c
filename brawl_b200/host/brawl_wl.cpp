// brawl_wl.cpp -- C++ host side of the Wang-Landau driver: wl_main (src/wang-landau.F90:101-314) on the C ABI.
//
// Same input files (brawl.inp, *.vij, wl_input.inp), same outputs (data/wl_dos_bins.nc, data/wl_dos.nc, data/wl_hist.nc,
// NetCDF-3 classic as ncdf_writer_1d writes them).  `ranks=N` plays `mpirun -np N`: N walkers, N / num_windows per window
// (:121-131).  Every window lives on one GPU; with gpus=G the windows are sharded over G processes (one per GPU) that talk
// through the C ABI's NCCL communicator (brawl_cuda_comm_*), exactly where the reference talks MPI:
//   sweeps + MPI_Allreduce/num_walkers (:539-631)  ->  brawl_cuda_wl_iterate (device-resident ln g / hist)
//   MPI_ALLREDUCE(converged) (:230), walker energies for replica_exchange (:1435)  ->  brawl_cuda_comm_allgather
//   dos_combine's send/recv + bcast (:1161-1192)  ->  brawl_cuda_wl_allgather_lng, stitched identically on every rank
//   replica_exchange's config swaps (:1485-1495)  ->  brawl_cuda_swap_replicas_batch / brawl_cuda_exchange_replicas
//   MPI_ALLREDUCE(wl_mc_steps) (:244)  ->  brawl_cuda_wl_allreduce
// The host arithmetic (divide_range, create_overlap, create_energy_bins, bin_index, dos_combine, replica_exchange's
// pairing and acceptance on the walkers' own MT19937 streams, mpi_window_optimise, compute_mean_energy) restates the
// reference line by line; tests/test_host_driver.py compares it with the oracle through the brawl_host_wl_* hooks.
// Not mirrored here: rho(E) sampling (radial_samples; brawl_b200/wang_landau.py has it), energy_explore / merge_configs
// (after a resize the walkers are steered into their new windows on the GPU instead of reloading stored configurations).
#include <unistd.h>
#include <sys/wait.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <limits>
#include <numeric>
#include <sstream>
#include <thread>

#include "brawl_host.hpp"

namespace brawl {

// ---- read_wl_file (src/io.f90:951-1086) ---------------------------------------------------------------------------
WLParams read_wl_file(const std::string &filename) {
  std::ifstream f(filename);
  if (!f) throw Stop("Could not parse wang landau input file. Aborting...");
  WLParams p;
  bool check[10] = {false, false, false, false, false, false, false, false, false, false};
  std::string line;
  auto first_token = [](const std::string &v) { std::istringstream s(v); std::string t; s >> t; return t; };
  while (std::getline(f, line)) {
    const size_t pos = line.find('=');
    if (pos == std::string::npos) continue;
    const std::string label = line.substr(0, pos);            // label = buffer(1:pos-1): leading blanks break a key there too
    const std::string t = first_token(line.substr(pos + 1));
    if (t.empty()) continue;
    std::string key = label;
    while (!key.empty() && key.back() == ' ') key.pop_back();
    if (key == "mc_sweeps") { p.mc_sweeps = std::atoi(t.c_str()); check[0] = true; }
    else if (key == "bins") { p.bins = std::atoi(t.c_str()); check[1] = true; }
    else if (key == "num_windows") { p.num_windows = std::atoi(t.c_str()); check[2] = true; }
    else if (key == "bin_overlap") { p.bin_overlap = std::strtof(t.c_str(), nullptr); check[3] = true; }
    else if (key == "tolerance") { p.tolerance = std::strtof(t.c_str(), nullptr); check[4] = true; }
    else if (key == "flatness") { p.flatness = std::strtof(t.c_str(), nullptr); check[5] = true; }
    else if (key == "wl_f") { p.wl_f = std::strtof(t.c_str(), nullptr); check[6] = true; }
    else if (key == "energy_min") { p.energy_min = std::strtof(t.c_str(), nullptr); check[7] = true; }
    else if (key == "energy_max") { p.energy_max = std::strtof(t.c_str(), nullptr); check[8] = true; }
    else if (key == "radial_samples") { p.radial_samples = std::atoi(t.c_str()); check[9] = true; }
    else if (key == "performance") p.performance = std::atoi(t.c_str());
    else if (key == "nbr_swap") { std::string u = t; for (auto &c : u) c = (char)std::toupper(c); p.nbr_swap = u.find('T') != std::string::npos && u.find('T') < 2; }
  }
  for (bool c : check) if (!c) throw Stop("Missing parameter in wang landau input file");
  return p;
}

// ---- window bookkeeping (1-based inclusive bin indices, as in the reference) ------------------------------------------
std::vector<int64_t> wl_divide_range(int bins, int W) {                       // divide_range (:855-891), power = 1
  std::vector<int64_t> iv(2 * (size_t)W, 0);
  iv[0] = 1; iv[2 * (W - 1) + 1] = bins;
  const double factor = ((double)bins - 1.0) / (((double)W + 1.0) - 1.0);
  for (int i = 2; i <= W; i++) {
    iv[2 * (i - 2) + 1] = (int64_t)std::floor(factor * (double)(i - 1) + 1.0);
    iv[2 * (i - 1)] = iv[2 * (i - 2) + 1] + 1;
  }
  return iv;
}
std::vector<int64_t> wl_create_overlap(const std::vector<int64_t> &iv, float bin_overlap) {   // create_overlap (:934-955)
  std::vector<int64_t> idx = iv;
  const int W = (int)iv.size() / 2;
  auto ext = [&](int64_t b) { return std::max<int64_t>((int64_t)std::ceil(bin_overlap * (float)b), 2); };   // single precision, like the reference
  if (W > 1) {
    for (int i = 2; i < W; i++) {
      const int64_t b = idx[2 * (i - 2) + 1] - idx[2 * (i - 2)] + 1;
      idx[2 * (i - 1)] = iv[2 * (i - 1)] - ext(b);
      idx[2 * (i - 1) + 1] = iv[2 * (i - 1) + 1];
    }
    const int64_t b = idx[2 * (W - 2) + 1] - idx[2 * (W - 2)];              // the last window's width lacks the +1 (reference quirk)
    idx[2 * (W - 1)] = iv[2 * (W - 1)] - ext(b);
    idx[2 * (W - 1) + 1] = iv[2 * (W - 1) + 1];
  }
  return idx;
}
std::vector<double> wl_create_energy_bins(int n_atoms, float energy_min, float energy_max, int bins, double *bin_width) {   // :969-986
  const double energy_to_ry = (double)n_atoms / (Ry_to_eV * 1000);
  const double width = (double)(energy_max - energy_min) / (double)(float)bins * energy_to_ry;
  std::vector<double> e((size_t)bins + 1);
  for (int i = 0; i <= bins; i++) e[i] = (double)energy_min * energy_to_ry + (double)i * width;
  if (bin_width) *bin_width = width;
  return e;
}
int wl_bin_index(double e, const std::vector<double> &edges, int bins) {                     // bin_index (:515-523)
  return (int)(((e - edges[0]) / (edges[bins] - edges[0])) * (double)bins) + 1;
}

// dos_combine (:1147-1194) as rank 0 evaluates it; lng[W][bins] window-averaged, win[W][2]
std::vector<double> wl_dos_combine(const std::vector<double> &lng, const std::vector<int64_t> &win, int W, int bins) {
  std::vector<double> comb(lng.begin(), lng.begin() + bins);
  int beta_index = 0;                                                        // not reset between windows, as in the reference
  for (int i = 2; i <= W; i++) {
    const double *buf = lng.data() + (size_t)(i - 1) * bins;
    const int start = (int)win[2 * (i - 1)], end = (int)win[2 * (i - 1) + 1];
    double beta_diff = std::numeric_limits<double>::max();
    const int jmax = (int)(win[2 * (i - 2) + 1] - win[2 * (i - 1)] - 1);
    for (int j = 0; j <= jmax; j++) {
      const double b_orig = comb[start + j] - comb[start + j - 1];
      const double b_merge = buf[start + j] - buf[start + j - 1];
      if (std::fabs(b_orig - b_merge) < beta_diff) { beta_diff = std::fabs(b_orig - b_merge); beta_index = start + j + 1; }
    }
    for (int j = beta_index; j <= end; j++) comb[j - 1] = buf[j - 1] + comb[beta_index - 1] - buf[beta_index - 1];
  }
  const double mn = *std::min_element(comb.begin(), comb.end());
  for (double &v : comb) v = v - mn;
  return comb;
}

// replica_exchange (:1392-1519): pairing on rank 0's stream (shuffle_rows), acceptance on the lower walker's stream.
// energies[P], lng[P][bins] per rank (P = W * walkers, rank = window-major), mts[P].  Returns the exchanges (lower, upper).
std::vector<std::pair<int, int>> wl_replica_exchange(const std::vector<double> &energies, const double *lng, const std::vector<int64_t> &win,
                                                     int W, int walkers, const std::vector<double> &edges, int bins, MT19937 *mts) {
  const int P = W * walkers;
  std::vector<int> ibin(P), loc(P);
  for (int r = 0; r < P; r++) {
    const int q = r / walkers + 1, ib = wl_bin_index(energies[r], edges, bins);
    ibin[r] = ib;
    bool lo = false, up = false;
    if (q > 1) lo = (ib < win[2 * (q - 2) + 1] + 1) && (ib > win[2 * (q - 1)] - 1);
    if (q < W) up = (ib > win[2 * q] - 1) && (ib < win[2 * (q - 1) + 1] + 1);
    loc[r] = up ? q : (lo ? q - 1 : 0);
  }
  auto shuffle_rows = [&](std::vector<std::pair<int, int>> &rows) {          // :1512-1519
    const int n = (int)rows.size();
    for (int i = n; i >= 2; i--) {
      int r = 1 + (int)(mts[0].genrand() * (double)i);
      if (r > n) r = n;
      std::swap(rows[i - 1], rows[r - 1]);
    }
  };
  std::vector<std::pair<int, int>> out;
  std::vector<char> accept(P, 0);
  for (int i = 1; i <= W - 1; i++) {
    std::vector<std::pair<int, int>> lower(walkers), upper(walkers), exch;
    for (int k = 0; k < walkers; k++) {
      lower[k] = {(i - 1) * walkers + k, loc[(i - 1) * walkers + k]};
      upper[k] = {i * walkers + k, loc[i * walkers + k]};
    }
    shuffle_rows(lower);
    shuffle_rows(upper);
    for (int j = 0; j < walkers; j++) {
      if (lower[j].second == 0) continue;
      for (int k = 0; k < walkers; k++) {
        if (upper[k].second == 0) continue;
        if (lower[j].second == upper[k].second) {
          exch.push_back({lower[j].first, upper[k].first});
          lower[j] = {0, 0}; upper[k] = {0, 0};
        }
      }
    }
    for (auto &ab : exch) {
      const int a = ab.first, b = ab.second;
      const int jb = wl_bin_index(energies[b], edges, bins);
      const double *la = lng + (size_t)a * bins;
      if (mts[a].genrand() < std::exp(la[ibin[a] - 1] - la[jb - 1])) accept[a] = 1;          // :1482-1483
      if (accept[a]) out.push_back({a, b});
    }
  }
  return out;
}

static double seq_sum(const std::vector<double> &v) { double t = 0.0; for (double x : v) t = t + x; return t; }
static std::vector<int> sort_descending(const std::vector<int64_t> &a) {     // :1686-1701 (exchange sort of an index vector)
  const int n = (int)a.size();
  std::vector<int> idx(n);
  std::iota(idx.begin(), idx.end(), 1);
  for (int i = 0; i < n - 1; i++)
    for (int j = i + 1; j < n; j++)
      if (a[idx[i] - 1] < a[idx[j] - 1]) std::swap(idx[i], idx[j]);
  return idx;
}
static int64_t nint(double x) { return x >= 0 ? (int64_t)std::floor(x + 0.5) : -(int64_t)std::floor(-x + 0.5); }

// the rank-0 arithmetic of mpi_window_optimise (:1211-1311); iv[W][2] and prev[W] are updated in place
void wl_window_optimise(int it, std::vector<int64_t> &iv, const std::vector<double> &mc_steps, std::vector<double> &prev, int bins) {
  const int W = (int)mc_steps.size();
  if (W < 2) return;
  double alpha = 0.8 * std::pow(0.8, (double)(it - 1));
  if (it == 0) alpha = 1.0;
  const double w_min = 0.02;
  std::vector<double> w_mc(W), frac(W);
  for (int i = 0; i < W; i++) {
    const int64_t first = iv[2 * i], last = iv[2 * i + 1];
    w_mc[i] = 1.0 / (mc_steps[i] / (double)(float)std::llabs(first - last + 1));      // width-2: x/0 = Inf, 1/Inf = 0 (IEEE)
  }
  double s = seq_sum(w_mc);
  for (double &x : w_mc) x = x / s;
  for (int i = 0; i < W; i++) frac[i] = alpha * w_mc[i] + (1.0 - alpha) * prev[i];
  s = seq_sum(frac);
  for (double &x : frac) x = x / s;
  prev = frac;
  std::vector<char> free_mask(W);
  for (int i = 0; i < W; i++) { frac[i] = std::max(frac[i], w_min); free_mask[i] = frac[i] > w_min; }
  if (std::fabs(seq_sum(frac) - 1.0) > 1.0e-12) {
    const double rem = 1.0 - seq_sum(frac);
    bool any = false;
    std::vector<double> fr;
    for (int i = 0; i < W; i++) if (free_mask[i]) { any = true; fr.push_back(frac[i]); }
    if (any) {
      const double sum_free = seq_sum(fr);
      if (sum_free > 0.0) {
        const double scale = rem / sum_free;
        for (int i = 0; i < W; i++) if (free_mask[i]) frac[i] = frac[i] + frac[i] * scale;
      }
    }
  }
  s = seq_sum(frac);
  for (double &x : frac) x = x / s;
  std::vector<int64_t> nb(W);
  const int64_t min_bins = std::max<int64_t>((int64_t)(w_min * bins), 2);
  for (int i = 0; i < W; i++) nb[i] = std::max(nint((double)(float)bins * frac[i]), min_bins);
  int64_t total = std::accumulate(nb.begin(), nb.end(), (int64_t)0);
  if (total != bins) {
    const std::vector<int> idx = sort_descending(nb);
    int64_t diff = bins - total;
    long i = 1, guard = 0;
    while (diff != 0) {
      const int j = idx[(i - 1) % W] - 1;
      if (diff > 0) { nb[j] += 1; diff -= 1; }
      else if (nb[j] > min_bins) { nb[j] -= 1; diff += 1; }
      i++;
      if (++guard > 4L * W * bins) throw Stop("window_optimise: the bins cannot hold the windows");
    }
  }
  iv[1] = nb[0];
  for (int i = 1; i < W; i++) { iv[2 * i] = iv[2 * (i - 1) + 1] + 1; iv[2 * i + 1] = iv[2 * i] + nb[i] - 1; }
  iv[2 * (W - 1)] = iv[2 * (W - 2) + 1] + 1;
  iv[2 * (W - 1) + 1] = bins;
}

// compute_mean_energy (:457-477): out[300][2] = (<E>(T), beta) at T = 10 .. 3000 K
std::vector<double> wl_compute_mean_energy(const std::vector<double> &lng, const std::vector<double> &edges, int bins, double bin_width) {
  std::vector<double> buf(bins), centre(bins), prob(bins), out(600);
  const double mx = *std::max_element(lng.begin(), lng.begin() + bins);
  for (int i = 0; i < bins; i++) { buf[i] = lng[i] - mx; centre[i] = edges[i] + 0.5 * bin_width; }
  for (int it = 1; it <= 300; it++) {
    const double beta = 1.0 / (k_b_in_Ry * (double)it * 10.0);
    double pm = -std::numeric_limits<double>::max();
    for (int i = 0; i < bins; i++) { prob[i] = buf[i] - beta * centre[i]; pm = std::max(pm, prob[i]); }
    for (int i = 0; i < bins; i++) prob[i] = std::exp(prob[i] - pm);
    const double s = seq_sum(prob);
    double e = 0.0;
    for (int i = 0; i < bins; i++) { prob[i] = prob[i] / s; e = e + centre[i] * prob[i]; }
    out[2 * (it - 1)] = e; out[2 * (it - 1) + 1] = beta;
  }
  return out;
}

// ---- wl_main ------------------------------------------------------------------------------------------------------
namespace {
struct WlRank {
  int rank = 0, world = 1;
  brawl_cuda_t *h = nullptr;
  std::vector<double> all_gather(const std::vector<double> &v) const {
    if (world == 1) return v;
    std::vector<double> out(v.size() * world);
    Gpu::check(brawl_cuda_comm_allgather(h, v.data(), (int)v.size(), out.data()));
    return out;
  }
  void all_reduce(std::vector<double> &v) const {
    if (world > 1) Gpu::check(brawl_cuda_wl_allreduce(h, v.data(), (int)v.size()));
  }
};

std::vector<double> run_wl_rank(RunParams &setup, const WLParams &p, const DriverOptions &opt, int rank, int world, const uint8_t *uid,
                                double *seconds, std::vector<int> *stage_sweeps) {
  const int W = p.num_windows, walkers = opt.ranks / W, bins = p.bins;
  const int w_local = W / world, first_window = rank * w_local, n_local = w_local * walkers, P = W * walkers;
  const std::vector<double> V = read_exchange(setup);
  Gpu gpu(setup, V, opt.device + rank, n_local);
  WlRank comm{rank, world, gpu.h};
  if (world > 1) Gpu::check(brawl_cuda_comm_create(gpu.h, world, rank, uid));
  double bin_width = 0.0;
  const std::vector<double> edges = wl_create_energy_bins(setup.n_atoms, p.energy_min, p.energy_max, bins, &bin_width);
  std::vector<int64_t> intervals = wl_divide_range(bins, W);
  std::vector<int64_t> win = wl_create_overlap(intervals, p.bin_overlap);
  Gpu::check(brawl_cuda_wl_init(gpu.h, bins, edges.data(), walkers));
  std::vector<int32_t> lo(n_local), hi(n_local);
  auto set_windows = [&]() {
    for (int w = 0; w < n_local; w++) { const int q = first_window + w / walkers; lo[w] = (int32_t)win[2 * q]; hi[w] = (int32_t)win[2 * q + 1]; }
    Gpu::check(brawl_cuda_wl_set_windows(gpu.h, lo.data(), hi.data(), 1));
  };
  set_windows();
  // one MT19937 stream per walker = per MPI rank of the reference (initialise_prng): identical on every process, so the
  // exchange plan is the same everywhere without a broadcast
  std::vector<MT19937> mts(P);
  for (int r = 0; r < P; r++) mts[r].f90_init_genrand(setup.static_seed ? 0 : 1, r, 0);
  // Philox key of this process's Monte-Carlo kernels (walker ids in the counters are handle-local)
  const uint64_t mc_seed = opt.seed ^ (0x9E3779B97F4A7C15ull * (uint64_t)(rank + 1));
  uint64_t offset = 0, rand_calls = 0;
  // species quotas exactly as initial_setup forms them (initialise.F90:468-506): count the species of one host-built start state
  std::vector<int64_t> counts(setup.n_species, 0);
  {
    MT19937 scratch;
    scratch.f90_init_genrand(0, 0, 0);
    Config c0;
    initial_setup(setup, c0, scratch);
    for (int8_t v : c0) if (v > 0) counts[v - 1]++;
  }
  std::vector<double> energies(n_local, 0.0);

  auto enter_energy_windows = [&](bool fresh) {                               // enter_energy_window (:643-741)
    std::vector<double> target(n_local), lo_e(n_local), hi_e(n_local);
    for (int w = 0; w < n_local; w++) {
      const double a = edges[lo[w] - 1], b = edges[hi[w]], cond = std::fabs(b - a) * 0.1;
      target[w] = (a + b) / 2.0; lo_e[w] = a + cond; hi_e[w] = b - cond;
    }
    float sf = 0.0025f * std::fabs(p.energy_max - p.energy_min);
    sf = sf * (float)setup.n_atoms;
    const double sigma = (double)sf / (Ry_to_eV * 1000);
    const double inv = 1.0 / (2.0 * sigma * sigma);
    auto rand_offset = [&]() { rand_calls++; return ((uint64_t)0x5A << 56) | ((uint64_t)rank << 32) | rand_calls; };
    if (fresh) Gpu::check(brawl_cuda_random_config(gpu.h, 0, n_local, counts.data(), opt.seed, rand_offset()));
    std::vector<int32_t> entered(n_local);
    for (int round = 0; round < 200; round++) {
      Gpu::check(brawl_cuda_wl_enter_window(gpu.h, n_local, target.data(), lo_e.data(), hi_e.data(), inv, (int64_t)setup.n_atoms * 250,
                                            mc_seed, offset++, energies.data(), entered.data()));
      bool all = true;
      const uint64_t off = rand_offset();
      for (int w = 0; w < n_local; w++)
        if (!entered[w]) { all = false; Gpu::check(brawl_cuda_random_config(gpu.h, w, 1, counts.data(), opt.seed, off)); }   // :671-674
      if (all) return;
    }
    throw Stop("walkers failed to enter their energy windows");
  };

  auto window_lng_all = [&]() {                                               // [W][bins]
    std::vector<double> all((size_t)W * bins);
    if (world > 1) Gpu::check(brawl_cuda_wl_allgather_lng(gpu.h, all.data()));
    else Gpu::check(brawl_cuda_wl_get(gpu.h, 0, all.data()));
    return all;
  };

  std::vector<double> wl_mc_steps(W, 0.0), diffusion_prev(W, 1.0 / (double)(float)W), hist_min(w_local), hist_mean(w_local);
  const int64_t n_trials = (int64_t)p.mc_sweeps * setup.n_atoms;
  std::vector<double> combined(bins, 0.0);

  auto stage = [&](double wl_f, double min_hist, int exchange_every, int it, bool resize) {
    std::vector<char> converged(w_local, 0);
    int n = 0;
    for (;;) {
      n++;
      Gpu::check(brawl_cuda_wl_iterate(gpu.h, wl_f, n_trials, p.nbr_swap ? 1 : 0, mc_seed, offset++, energies.data(), hist_min.data(),
                                       hist_mean.data(), nullptr));
      int n_conv = 0;
      for (int q = 0; q < w_local; q++) {
        if (!converged[q]) wl_mc_steps[first_window + q] += (double)n_trials * walkers;      // every walker adds its trials (:217)
        const double flat = hist_mean[q] > 0.0 ? hist_min[q] / hist_mean[q] : 0.0;           // :222
        bool good = flat > (double)p.flatness;
        if (min_hist >= 0.0) good = good && hist_min[q] > min_hist;                          // pre_sampling (:781-783)
        if (good) converged[q] = 1;
        n_conv += converged[q];
      }
      std::vector<double> send(energies);
      send.push_back((double)n_conv);
      const std::vector<double> both = comm.all_gather(send);                 // energies (:1435) + converged count (:230)
      std::vector<double> e_all((size_t)P);
      double conv_sum = 0.0;
      for (int r = 0; r < world; r++) {
        std::copy(both.begin() + (size_t)r * (n_local + 1), both.begin() + (size_t)r * (n_local + 1) + n_local, e_all.begin() + (size_t)r * n_local);
        conv_sum += both[(size_t)r * (n_local + 1) + n_local];
      }
      if (n % exchange_every == 0 && W > 1 && (p.performance == 0 || p.performance == 2 || p.performance == 4)) {
        const std::vector<double> lw = window_lng_all();
        std::vector<double> lng_ranks((size_t)P * bins);                      // each rank's wl_logdos = its window's average
        for (int r = 0; r < P; r++) std::copy(lw.begin() + (size_t)(r / walkers) * bins, lw.begin() + (size_t)(r / walkers + 1) * bins, lng_ranks.begin() + (size_t)r * bins);
        const auto swaps = wl_replica_exchange(e_all, lng_ranks.data(), win, W, walkers, edges, bins, mts.data());
        std::vector<int32_t> la, lb, rr, rp;
        for (auto &ab : swaps) {
          const int ra = ab.first / n_local, rb = ab.second / n_local, a = ab.first % n_local, b = ab.second % n_local;
          if (ra == rank && rb == rank) { la.push_back(a); lb.push_back(b); std::swap(energies[a], energies[b]); }
          else if (ra == rank) { rr.push_back(a); rp.push_back(rb); energies[a] = e_all[ab.second]; }
          else if (rb == rank) { rr.push_back(b); rp.push_back(ra); energies[b] = e_all[ab.first]; }
        }
        if (!la.empty()) Gpu::check(brawl_cuda_swap_replicas_batch(gpu.h, (int)la.size(), la.data(), lb.data()));
        if (!rr.empty()) Gpu::check(brawl_cuda_exchange_replicas(gpu.h, (int)rr.size(), rr.data(), rp.data()));
      }
      if ((int)conv_sum == W) break;
    }
    if (stage_sweeps) stage_sweeps->push_back(n);
    combined = wl_dos_combine(window_lng_all(), win, W, bins);                // dos_average + dos_combine (:240-241)
    Gpu::check(brawl_cuda_wl_set_lng(gpu.h, combined.data()));
    Gpu::check(brawl_cuda_wl_zero_hist(gpu.h));
    comm.all_reduce(wl_mc_steps);                                             // :244
    if (resize && W > 1) {                                                    // mpi_window_optimise (:284-286): same arithmetic on every rank
      wl_window_optimise(it, intervals, wl_mc_steps, diffusion_prev, bins);
      win = wl_create_overlap(intervals, p.bin_overlap);
      set_windows();
    }
    std::fill(wl_mc_steps.begin(), wl_mc_steps.end(), 0.0);
    if (resize && W > 1) enter_energy_windows(false);                         // in place of load_window_config
  };

  Gpu::check(brawl_cuda_synchronize(gpu.h));
  const auto t0 = std::chrono::steady_clock::now();
  enter_energy_windows(true);
  double wl_f = (double)p.wl_f;
  stage(wl_f, 1000.0 / (double)(float)walkers, 10, 0, p.performance >= 0 && p.performance <= 3);   // pre_sampling (:757-838)
  int it = 1;
  while (wl_f > (double)p.tolerance) {                                        // :198-292
    stage(wl_f, -1.0, 1, it, p.performance == 0 || p.performance == 1);
    it++;
    wl_f = wl_f * 0.5;
  }
  Gpu::check(brawl_cuda_synchronize(gpu.h));
  if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (rank == 0) {                                                            // save_wl_data (:409-417)
    mkdir_p("data");
    ncdf_writer_1d("data/wl_dos_bins.nc", edges);
    ncdf_writer_1d("data/wl_dos.nc", combined);
    ncdf_writer_1d("data/wl_hist.nc", std::vector<double>(bins, 0.0));
  }
  return combined;
}
}  // namespace

void wl_main(RunParams &setup, const WLParams &p, const DriverOptions &opt) {
  const int W = p.num_windows, G = std::max(1, opt.gpus);
  std::printf(" MPI processes: %d\n", opt.ranks);
  if (W < 1 || opt.ranks % W != 0) {                                          // :121-131 (the reference exits with status 0)
    std::printf("%s\n~~~~~ Error: Number of MPI processes not divisible by num_windows ~~~~~~\n%s\n", std::string(72, '~').c_str(), std::string(72, '~').c_str());
    return;
  }
  if (W % G != 0) throw Stop("num_windows must be divisible by the number of GPUs");
  double seconds = 0.0;
  std::vector<int> sweeps;
  if (G == 1) {
    run_wl_rank(setup, p, opt, 0, 1, nullptr, &seconds, &sweeps);
  } else {
    // one process per GPU (fork before any CUDA call); rank 0 publishes the NCCL id through a file in the run directory
    const std::string idfile = ".brawl_nccl_id";
    std::remove(idfile.c_str());
    std::fflush(stdout);
    std::vector<pid_t> kids;
    int rank = 0;
    for (int g = 1; g < G; g++) {
      pid_t pid = fork();
      if (pid < 0) throw Stop("fork failed");
      if (pid == 0) { rank = g; kids.clear(); break; }
      kids.push_back(pid);
    }
    uint8_t uid[128];
    if (rank == 0) {
      Gpu::check(brawl_cuda_comm_unique_id(uid));
      std::ofstream f(idfile + ".tmp", std::ios::binary);
      f.write((const char *)uid, 128);
      f.close();
      std::rename((idfile + ".tmp").c_str(), idfile.c_str());
    } else {
      for (int tries = 0;; tries++) {
        std::ifstream f(idfile, std::ios::binary);
        if (f && f.read((char *)uid, 128)) break;
        if (tries > 6000) throw Stop("timed out waiting for the NCCL id of rank 0");
        std::this_thread::sleep_for(std::chrono::milliseconds(10));
      }
    }
    int status = 0;
    try {
      run_wl_rank(setup, p, opt, rank, G, uid, &seconds, &sweeps);
    } catch (const Stop &e) {
      std::fprintf(stderr, "STOP (rank %d) %s\n", rank, e.what());
      status = 1;
    }
    if (rank != 0) _exit(status);
    for (pid_t k : kids) { int st = 0; waitpid(k, &st, 0); if (!WIFEXITED(st) || WEXITSTATUS(st)) status = 1; }
    std::remove(idfile.c_str());
    if (status) throw Stop("a Wang-Landau rank failed");
  }
  std::printf(" Simulation Complete!  %d windows x %d walkers on %d GPU(s): %.3f s to the final ln g(E); sweeps calls per stage:", W, opt.ranks / W, G, seconds);
  for (int n : sweeps) std::printf(" %d", n);
  std::printf("\n");
}

}  // namespace brawl

// ---- C hooks for the parity tests (ctypes): the host arithmetic against the oracle ------------------------------------
extern "C" {
void brawl_host_wl_divide_range(int bins, int W, float bin_overlap, int64_t *intervals, int64_t *window_indices) {
  auto iv = brawl::wl_divide_range(bins, W);
  auto idx = brawl::wl_create_overlap(iv, bin_overlap);
  std::copy(iv.begin(), iv.end(), intervals);
  std::copy(idx.begin(), idx.end(), window_indices);
}
void brawl_host_wl_energy_bins(int n_atoms, float e_min, float e_max, int bins, double *edges) {
  auto e = brawl::wl_create_energy_bins(n_atoms, e_min, e_max, bins, nullptr);
  std::copy(e.begin(), e.end(), edges);
}
void brawl_host_wl_dos_combine(const double *lng, const int64_t *win, int W, int bins, double *out) {
  std::vector<double> l(lng, lng + (size_t)W * bins);
  std::vector<int64_t> w(win, win + 2 * (size_t)W);
  auto c = brawl::wl_dos_combine(l, w, W, bins);
  std::copy(c.begin(), c.end(), out);
}
int brawl_host_wl_replica_exchange(const double *energies, const double *lng, const int64_t *win, int W, int walkers, const double *edges,
                                   int bins, uint32_t *mt_states625, int *pairs) {
  const int P = W * walkers;
  std::vector<brawl::MT19937> mts(P);
  for (int r = 0; r < P; r++) mts[r].import625(mt_states625 + (size_t)r * 625);
  std::vector<double> e(energies, energies + P), ed(edges, edges + bins + 1);
  std::vector<int64_t> w(win, win + 2 * (size_t)W);
  auto sw = brawl::wl_replica_exchange(e, lng, w, W, walkers, ed, bins, mts.data());
  for (int r = 0; r < P; r++) mts[r].export625(mt_states625 + (size_t)r * 625);
  for (size_t i = 0; i < sw.size(); i++) { pairs[2 * i] = sw[i].first; pairs[2 * i + 1] = sw[i].second; }
  return (int)sw.size();
}
void brawl_host_wl_window_optimise(int it, int W, int64_t *iv, const double *mc_steps, double *prev, int bins) {
  std::vector<int64_t> v(iv, iv + 2 * (size_t)W);
  std::vector<double> m(mc_steps, mc_steps + W), p(prev, prev + W);
  brawl::wl_window_optimise(it, v, m, p, bins);
  std::copy(v.begin(), v.end(), iv);
  std::copy(p.begin(), p.end(), prev);
}
void brawl_host_wl_mean_energy(const double *lng, const double *edges, int bins, double bin_width, double *out600) {
  std::vector<double> l(lng, lng + bins), e(edges, edges + bins + 1);
  auto o = brawl::wl_compute_mean_energy(l, e, bins, bin_width);
  std::copy(o.begin(), o.end(), out600);
}
}
