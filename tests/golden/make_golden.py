#!/usr/bin/env python3
"""Build tests/golden/brawl_golden.npz from the reference's own regression fixtures.

/root/reference does not exist on the GPU box, so the golden vectors that pin the oracle
 are converted ONCE, here in the build container, into one small
compressed archive that is committed next to this script.  Sources (all under
/root/reference/tests):

  01_serial_metropolis/{brawl.inp,metropolis.inp,fcc_epi.vij}         + 99_ref/01_serial_metropolis/*
  02_parallel_metropolis/{...,bcc_epi.vij}                            + 99_ref/02_parallel_metropolis/*
  03_serial_nested_sampling/{...,fcc_al_1.00_crfeconi.vij}            + 99_ref/03_serial_nested_sampling/*
  04_parallel_wang-landau/{...,bcc_epi.vij}                           + 99_ref/04_parallel_wang-landau/wl_dos.nc
  examples/02_wang-landau_AlTiCrMo/AlTiCrMo.vij, examples/01_metropolis_FeNi/.../FeNi.vij,
  examples/03_nested_sampling_AlCrFeCoNi/*.vij   (interaction tables used by bench.py's synthetic workloads)

Text goldens are stored verbatim (the reference's comparator demands identical lines,
tests/ci_test.py:61-69); NetCDF goldens are stored as their arrays (ci_test.py:23-54 compares
arrays only).  Run:  python tests/golden/make_golden.py
"""
import glob
import os

import numpy as np
from scipy.io import netcdf_file

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def vij(path):
    # read(16,*) V_ex : list-directed read, fills V_ex(S,S,n) first index fastest (src/io.f90:401)
    return np.array(open(path).read().split(), dtype=np.float64)


def nc_vars(path):
    nc = netcdf_file(path, "r", mmap=False)
    return {k: np.array(v.data) for k, v in nc.variables.items()}


def text(path):
    return np.array(open(path).read())


def main():
    out = {}
    t = REF + "/tests/"
    r = t + "99_ref/"
    out["t01_V"] = vij(t + "01_serial_metropolis/fcc_epi.vij")
    out["t02_V"] = vij(t + "02_parallel_metropolis/bcc_epi.vij")
    out["t03_V"] = vij(t + "03_serial_nested_sampling/fcc_al_1.00_crfeconi.vij")
    out["t04_V"] = vij(t + "04_parallel_wang-landau/bcc_epi.vij")
    # --- 01
    d = r + "01_serial_metropolis/"
    out["t01_initial"] = nc_vars(d + "proc_0000_initial_config_at_0300.0.nc")["configuration"][..., 0].astype(np.int8)
    out["t01_final"] = nc_vars(d + "proc_0000_final_config_at_0300.0.nc")["configuration"][..., 0].astype(np.int8)
    out["t01_energy_txt"] = text(d + "proc_0000_energy_trajectory_at_T_0300.0.dat")
    out["t01_asro_txt"] = text(d + "proc_0000_asro_trajectory_at_T_0300.0.dat")
    for k, v in nc_vars(d + "proc_0000_rho_of_T.nc").items():
        out["t01_rho_" + k.split()[0]] = v
    # --- 02
    d = r + "02_parallel_metropolis/"
    for rank in range(4):
        p = "t02_r%d_" % rank
        f = d + "proc_%04d_" % rank
        out[p + "initial"] = nc_vars(f + "initial_config_at_0300.0.nc")["configuration"][..., 0].astype(np.int8)
        out[p + "final"] = nc_vars(f + "final_config_at_0300.0.nc")["configuration"][..., 0].astype(np.int8)
        out[p + "energy_txt"] = text(f + "energy_trajectory_at_T_0300.0.dat")
        out[p + "asro_txt"] = text(f + "asro_trajectory_at_T_0300.0.dat")
        out[p + "diag_txt"] = text(f + "energy_diagnostics.dat")
        for k, v in nc_vars(f + "rho_of_T.nc").items():
            out[p + "rho_" + k.split()[0]] = v
    out["t02_av_diag_txt"] = text(d + "av_energy_diagnostics.dat")
    for k, v in nc_vars(d + "av_radial_density.nc").items():
        out["t02_av_rho_" + k.split()[0]] = v
    # --- 03
    out["t03_energies_txt"] = text(r + "03_serial_nested_sampling/fcc_al_1.00_crfeconi_K100.energies")
    # --- 04
    out["t04_wl_dos"] = nc_vars(r + "04_parallel_wang-landau/wl_dos.nc")["grid data"]
    # --- interaction tables for the BASELINE.json workloads
    out["ex_AlTiCrMo_V"] = vij(REF + "/examples/02_wang-landau_AlTiCrMo/AlTiCrMo.vij")
    out["ex_FeNi_V"] = vij(glob.glob(REF + "/examples/01_metropolis_FeNi/**/FeNi.vij", recursive=True)[0])
    out["ex_AlCrFeCoNi_V"] = vij(glob.glob(REF + "/examples/03_nested_sampling_AlCrFeCoNi/**/*.vij", recursive=True)[0])
    # --- the reference's own input files for its regression cases (text) + two raw NetCDF goldens
    for case, files in (("01_serial_metropolis", ("brawl.inp", "metropolis.inp", "fcc_epi.vij")),
                        ("02_parallel_metropolis", ("brawl.inp", "metropolis.inp", "bcc_epi.vij")),
                        ("03_serial_nested_sampling", ("brawl.inp", "ns_input.inp", "fcc_al_1.00_crfeconi.vij")),
                        ("04_parallel_wang-landau", ("brawl.inp", "wl_input.inp", "bcc_epi.vij"))):
        for fn in files:
            out["in_%s_%s" % (case[:2], fn)] = text(t + case + "/" + fn)
    out["raw_t02_r0_initial_nc"] = np.frombuffer(open(r + "02_parallel_metropolis/proc_0000_initial_config_at_0300.0.nc", "rb").read(), dtype=np.uint8)
    out["raw_t02_r0_rho_nc"] = np.frombuffer(open(r + "02_parallel_metropolis/proc_0000_rho_of_T.nc", "rb").read(), dtype=np.uint8)
    out["raw_t04_wl_dos_nc"] = np.frombuffer(open(r + "04_parallel_wang-landau/wl_dos.nc", "rb").read(), dtype=np.uint8)
    out["raw_t02_av_rho_nc"] = np.frombuffer(open(r + "02_parallel_metropolis/av_radial_density.nc", "rb").read(), dtype=np.uint8)
    path = os.path.join(HERE, "brawl_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "entries")
    for k, v in out.items():
        print("  %-24s %s %s" % (k, v.dtype, v.shape))


if __name__ == "__main__":
    main()
