// brawl_host.hpp -- C++ host side above the C ABI: a Fortran-free mirror of the BraWl drivers for the
// swap hot path.  Same inputs (brawl.inp, metropolis.inp, ns_input.inp, *.vij), same outputs
// (NetCDF-3 classic configs / rho_of_T, text trajectories and diagnostics, .energies), same names and
// error texts as the reference routines cited at each declaration; every trial, energy and pair count
// runs on the GPU through include/brawl_cuda.h.  Product code: no reference to oracle/.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/brawl_cuda.h"

namespace brawl {

// the reference `stop`s with a message after comms_finalise(); here: an exception carrying that message
struct Stop : std::runtime_error { using std::runtime_error::runtime_error; };

// src/constants.f90:41-49 (k_b_in_eV carries the reference's digit transposition)
constexpr double k_b_in_Ry = 8.167333262e-5 / 13.605693122990;
constexpr double Ry_to_eV = 13.605693122;

// ---- type(run_params), src/derived_types.f90:56-100 ----------------------------------------------
struct RunParams {
  int mode = 301;
  int n_1 = 4, n_2 = 4, n_3 = 4, n_basis = 1, n_species = 4, n_atoms = 0;
  bool static_seed = false;
  std::string lattice = "fcc";
  double lattice_parameter = 3.57;
  std::vector<std::string> species_names;
  std::vector<double> species_concentrations;   // (0:n_species), index 0 = 0.0
  std::vector<int64_t> species_numbers;
  std::string interaction_file = "V_ijs.txt";
  int interaction_range = 0, wc_range = 2;
  int lattice_id() const;                        // 0 simple_cubic, 1 bcc, 2 fcc; throws like initialise.F90:238-240
};
// ---- type(metropolis_params), src/derived_types.f90:141-226 ---------------------------------------
struct MetropolisParams {
  std::string mode;
  int64_t n_mc_steps = 0, n_burn_in_steps = 0, n_sample_steps = 0;
  bool burn_in_start = false, burn_in = false, calculate_energies = true, write_trajectory_energy = false;
  bool calculate_asro = true, write_trajectory_asro = false, calculate_alro = false, write_trajectory_xyz = false;
  int64_t n_sample_steps_asro = 0, n_sample_steps_alro = 0, n_sample_steps_trajectory = 0;
  bool write_initial_config_xyz = false, write_initial_config_nc = false, write_final_config_xyz = false;
  bool write_final_config_nc = false, read_start_config_nc = false, nbr_swap = false;
  std::string start_config_file;
  double T = 0.0, delta_T = 1.0;
  int T_steps = 1;
};
// ---- type(ns_params), src/derived_types.f90:242-255 ------------------------------------------------
struct NSParams {
  int n_walkers = 0, n_steps = 0, n_iter = 0, traj_freq = 100;
  std::string outfile_ener, outfile_traj;
};

// ---- type(wl_params), src/derived_types.f90:322-350 (reals are single precision there) ----------------
struct WLParams {
  int mc_sweeps = 100, bins = 512, num_windows = 4, radial_samples = 8, performance = 0;
  float bin_overlap = 0.25f, tolerance = 5e-5f, flatness = 0.9f, wl_f = 0.05f, energy_min = -96.0f, energy_max = 0.0f;
  bool nbr_swap = false;
};

// ---- parsers (src/io.f90) ---------------------------------------------------------------------------
RunParams read_control_file(const std::string &filename);               // io.f90:135-330
MetropolisParams read_metropolis_file(const std::string &filename);     // io.f90:519-686
NSParams read_ns_file(const std::string &filename);                     // io.f90:740-817
WLParams read_wl_file(const std::string &filename);                     // io.f90:951-1086
std::vector<double> read_exchange(const RunParams &setup);              // io.f90:389-415 -> V_ex(S,S,n_shells)
void initialise_function_pointers(RunParams &setup);                    // initialise.F90:153-257 (n_atoms, validation)

// ---- MT19937 (published algorithm; reference vendors it as src/mt19937ar.c) ----------------------------
struct MT19937 {
  uint32_t mt[624];
  int mti = 625;
  void init_genrand(uint32_t s);                                        // mt19937ar.c:63-78
  uint32_t genrand_int32();                                             // :82-118
  double genrand() { return genrand_int32() * (1.0 / 4294967296.0); }   // :121-125
  uint32_t f90_init_genrand(int seedtime, int my_rank, unsigned long job_id);   // :127-142
  void export625(uint32_t *s) const;
  void import625(const uint32_t *s);
};

// ---- configuration helpers ----------------------------------------------------------------------------
using Config = std::vector<int8_t>;                                     // grid[z][y][x], reference layout
void initial_setup(RunParams &setup, Config &config, MT19937 &rng);     // initialise.F90:434-617
std::vector<double> lattice_shells(const RunParams &setup, const Config &config);   // analytics.f90:205-275

// ---- writers ---------------------------------------------------------------------------------------------
void energy_trajectory_writer(const std::string &f, int64_t step, double energy);            // metropolis_output.f90:34-53
void asro_trajectory_writer(const std::string &f, int64_t step, const std::vector<double> &asro);   // :66-98 (asro in (i,j,k) Fortran order)
void diagnostics_writer(const std::string &f, const std::vector<double> &T, const std::vector<double> &E,
                        const std::vector<double> &C, const std::vector<double> &acc);       // :113-135
void ncdf_grid_state_writer(const std::string &f, const Config &state, const RunParams &setup);   // netcdf_io.f90:495-583
void ncdf_order_writer(const std::string &f, const std::vector<double> &order /* (species,basis,x,y,z,T) Fortran order */,
                       const std::vector<double> &T, const RunParams &s);
void ncdf_radial_density_writer(const std::string &f, const std::vector<double> &rho /* (i,j,r,T) Fortran order */,
                                const std::vector<double> &r, const std::vector<double> &T,
                                const std::vector<double> &U, const RunParams &setup);         // netcdf_io.f90:150-244
void ncdf_config_reader(const std::string &f, Config &config, const RunParams &setup);        // netcdf_io.f90:1368-1429
void ncdf_writer_1d(const std::string &f, const std::vector<double> &grid_data);              // netcdf_io.f90:731-806
void xyz_writer(const std::string &f, const Config &config, const RunParams &setup, bool trajectory = false);   // write_xyz.f90:39-116
void mkdir_p(const std::string &d);

// ---- GPU handle (RAII over brawl_cuda_t) --------------------------------------------------------------
struct Gpu {
  brawl_cuda_t *h = nullptr;
  Gpu(const RunParams &setup, const std::vector<double> &V, int device, int n_replicas);
  ~Gpu();
  Gpu(const Gpu &) = delete;
  static void check(int rc);
};

// ---- drivers ------------------------------------------------------------------------------------------------
struct DriverOptions {
  std::string rng = "mt19937";   // "mt19937": replay the reference stream bit-for-bit; "philox": production kernels
  int ranks = 1;                 // emulate `mpirun -np ranks`: rank r uses seed 110179+11 r and writes proc_000r files
  int device = 0;
  uint64_t seed = 0x42726157ull;
  int gpus = 1;                  // Wang-Landau: processes (one per GPU, devices device .. device+gpus-1) the windows are sharded over
};
void metropolis_main(RunParams &setup, MetropolisParams &metropolis, const DriverOptions &opt);   // metropolis.F90:46-71
void nested_sampling_main(RunParams &setup, const DriverOptions &opt);                             // nested_sampling.f90:45-206
void wl_main(RunParams &setup, const WLParams &wl_setup, const DriverOptions &opt);                // wang-landau.F90:101-314

// ---- Wang-Landau host arithmetic (1-based inclusive bin indices; src/wang-landau.F90 lines at each definition) ------
std::vector<int64_t> wl_divide_range(int bins, int W);
std::vector<int64_t> wl_create_overlap(const std::vector<int64_t> &intervals, float bin_overlap);
std::vector<double> wl_create_energy_bins(int n_atoms, float energy_min, float energy_max, int bins, double *bin_width);
int wl_bin_index(double e, const std::vector<double> &edges, int bins);
std::vector<double> wl_dos_combine(const std::vector<double> &lng, const std::vector<int64_t> &win, int W, int bins);
std::vector<std::pair<int, int>> wl_replica_exchange(const std::vector<double> &energies, const double *lng, const std::vector<int64_t> &win,
                                                     int W, int walkers, const std::vector<double> &edges, int bins, MT19937 *mts);
void wl_window_optimise(int it, std::vector<int64_t> &iv, const std::vector<double> &mc_steps, std::vector<double> &prev, int bins);
std::vector<double> wl_compute_mean_energy(const std::vector<double> &lng, const std::vector<double> &edges, int bins, double bin_width);

}  // namespace brawl
