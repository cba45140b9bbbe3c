// brawl_driver -- Fortran-free equivalent of `brawl.run` (src/main.F90:10-177) for the swap hot path:
// reads the reference's own input files from the current directory and writes its output files, with
// every trial / energy / pair count on the GPU (libbrawl_cuda.so).
//   brawl_driver [input=brawl.inp] [metropolis=metropolis.inp] [wl=wl_input.inp] [rng=mt19937|philox] [ranks=N] [gpus=G]
//                [device=D] [seed=S]
// rng=mt19937 (default) replays the reference's MT19937 stream: outputs are the reference's, bit for bit.
// ranks=N emulates `mpirun -np N` (independent replicas, seeds 110179+11*rank, proc_000r_* + av_* files; Wang-Landau:
// N walkers = N / num_windows per window, sharded with gpus=G over G processes / GPUs that talk NCCL through the C ABI).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "brawl_host.hpp"

int main(int argc, char **argv) {
  std::string input = "brawl.inp", metro = "metropolis.inp", wl = "wl_input.inp";
  brawl::DriverOptions opt;
  int dryrun = 0;   // dryrun=1: parse the inputs, build rank 0's initial configuration on the host and write it; no GPU
  for (int i = 1; i < argc; i++) {                       // key=value, like src/command_line.f90
    std::string a = argv[i];
    size_t p = a.find('=');
    if (p == std::string::npos) { std::fprintf(stderr, "ignoring argument '%s' (expected key=value)\n", a.c_str()); continue; }
    std::string k = a.substr(0, p), v = a.substr(p + 1);
    if (k == "input") input = v;
    else if (k == "metropolis") metro = v;
    else if (k == "wl") wl = v;
    else if (k == "gpus") opt.gpus = std::atoi(v.c_str());
    else if (k == "rng") opt.rng = v;
    else if (k == "ranks") opt.ranks = std::atoi(v.c_str());
    else if (k == "device") opt.device = std::atoi(v.c_str());
    else if (k == "seed") opt.seed = std::strtoull(v.c_str(), nullptr, 0);
    else if (k == "dryrun") dryrun = std::atoi(v.c_str());
  }
  try {
    brawl::RunParams setup = brawl::read_control_file(input);
    brawl::initialise_function_pointers(setup);
    if (dryrun) {
      std::printf("mode=%d lattice=%s n=%d,%d,%d n_species=%d n_atoms=%d interaction_file=%s interaction_range=%d wc_range=%d static_seed=%d\n",
                  setup.mode, setup.lattice.c_str(), setup.n_1, setup.n_2, setup.n_3, setup.n_species, setup.n_atoms,
                  setup.interaction_file.c_str(), setup.interaction_range, setup.wc_range, (int)setup.static_seed);
      std::vector<double> V = brawl::read_exchange(setup);
      std::printf("V_ex entries=%zu first=%.17g last=%.17g\n", V.size(), V.front(), V.back());
      if (setup.mode == 301) {
        brawl::MetropolisParams mp = brawl::read_metropolis_file(metro);
        std::printf("metropolis mode=%s n_mc_steps=%lld n_sample_steps=%lld asro=%lld alro=%lld traj=%lld T=%.6f T_steps=%d delta_T=%.6f burn_in=%d burn_in_start=%d nbr_swap=%d\n",
                    mp.mode.c_str(), (long long)mp.n_mc_steps, (long long)mp.n_sample_steps, (long long)mp.n_sample_steps_asro,
                    (long long)mp.n_sample_steps_alro, (long long)mp.n_sample_steps_trajectory, mp.T, mp.T_steps, mp.delta_T,
                    (int)mp.burn_in, (int)mp.burn_in_start, (int)mp.nbr_swap);
      }
      brawl::MT19937 rng;
      rng.f90_init_genrand(setup.static_seed ? 0 : 1, 0, 0);
      brawl::Config config;
      brawl::initial_setup(setup, config, rng);
      auto shells = brawl::lattice_shells(setup, config);
      std::printf("shells");
      for (double sh : shells) std::printf(" %.17g", sh);
      std::printf("\nnext_genrand=%.17g\n", rng.genrand());
      brawl::mkdir_p("configs");
      brawl::ncdf_grid_state_writer("configs/dryrun_initial_config.nc", config, setup);
      return 0;
    }
    if (setup.mode == 301) {                             // main.F90:89-100
      brawl::MetropolisParams mp = brawl::read_metropolis_file(metro);
      brawl::metropolis_main(setup, mp, opt);
    } else if (setup.mode == 303) {                      // main.F90:121-130 (serial only)
      if (opt.ranks > 1) throw brawl::Stop("Nested sampling is serial only");
      brawl::nested_sampling_main(setup, opt);
    } else if (setup.mode == 302) {                      // main.F90:101-120 (MPI build only): ranks = walkers in total
      brawl::WLParams wp = brawl::read_wl_file(wl);
      brawl::wl_main(setup, wp, opt);
    } else if (setup.mode == 304) {
      throw brawl::Stop("TMMC is WIP in the reference and does not function (src/tmmc.F90:67-73)");
    } else {
      throw brawl::Stop("Unrecognised mode");
    }
  } catch (const brawl::Stop &e) {
    std::fprintf(stderr, "STOP %s\n", e.what());
    return 1;
  }
  return 0;
}
