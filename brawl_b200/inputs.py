"""The reference's own input files for the Python drivers: brawl.inp (read_control_file, src/io.f90:150-330), *.vij
(read_exchange, src/io.f90:380-410) and the species quotas initial_setup derives (src/initialise.F90:468-506).  Host-side
text parsing only; the C++ host (brawl_b200/host) has the same parsers for brawl_driver."""
import math
import os

import numpy as np

from ._lib import BrawlCudaError

_REQUIRED = ("mode", "lattice", "lattice_parameter", "n_1", "n_2", "n_3", "n_species", "interaction_file", "interaction_range",
             "species_names")


def _logical(v):
    return v.strip().strip(".").upper().startswith("T")


def _string(v):
    v = v.strip()
    if v[:1] in "'\"":
        q = v[0]
        return v[1:v.find(q, 1) if v.find(q, 1) > 0 else None]
    return v.split()[0] if v.split() else ""


def read_control_file(path):
    """brawl.inp -> dict.  key = value lines, '#' comments, unknown keys ignored, missing mandatory keys are the
    reference's error messages ("Missing '<key>' in system file")."""
    if not os.path.exists(path):
        raise BrawlCudaError("Could not find input file " + path)
    kv = []
    for line in open(path):
        if line.lstrip().startswith("#") or "=" not in line:
            continue
        k, v = line.split("=", 1)
        kv.append((k.rstrip(" \t"), v.split("#")[0] if not v.strip().startswith(("'", '"')) else v))
    p = dict(wc_range=2, static_seed=False, n_basis=1)
    conv = dict(mode=int, lattice=_string, lattice_parameter=float, n_1=int, n_2=int, n_3=int, n_species=int,
                interaction_file=_string, interaction_range=int, wc_range=int, static_seed=_logical)
    for k, v in kv:
        if k in conv:
            p[k] = conv[k](v.strip()) if conv[k] in (int, float) else conv[k](v)
    S = p.get("n_species", 0)
    for k, v in kv:                                     # second pass: the species arrays need n_species (io.f90:249-289)
        t = v.replace(",", " ").split()
        if k == "species_names":
            p[k] = [x.strip("'\"")[:2] for x in t[:S]]
        elif k == "species_concentrations":
            p[k] = [0.0] + [float(x) for x in t[:S]]    # species_concentrations(0:) with element 0 = 0 (io.f90:242-246)
        elif k == "species_numbers":
            p[k] = [int(x) for x in t[:S]]
    if "species_concentrations" in p and "species_numbers" in p:
        raise BrawlCudaError("You cannot specify both chemical concentrations and numbers of atoms!")
    if "species_concentrations" not in p and "species_numbers" not in p:
        raise BrawlCudaError("You must specify either chemical concentrations or numbers of atoms.")
    for k in _REQUIRED:
        if k not in p:
            raise BrawlCudaError("Missing '%s' in system file" % k)
    if p["lattice"] not in ("simple_cubic", "bcc", "fcc"):
        raise BrawlCudaError("Lattice type not yet implemented!")
    return p


def read_exchange(path, n_species, n_shells):
    """*.vij: list-directed read of V_ex(S, S, n_shells), first index fastest (io.f90:401) -- the flat array Device takes."""
    if not os.path.exists(path):
        raise BrawlCudaError("Could not find interaction file " + path)
    v = np.array(open(path).read().replace(",", " ").split(), dtype=np.float64)
    need = n_species * n_species * n_shells
    if v.size < need:
        raise BrawlCudaError("%s holds %d numbers, %d x %d x %d needed" % (path, v.size, n_species, n_species, n_shells))
    return np.ascontiguousarray(v[:need])


def n_atoms(p):
    """initialise.F90:162,175,214."""
    return {"simple_cubic": 8, "bcc": 2, "fcc": 4}[p["lattice"]] * p["n_1"] * p["n_2"] * p["n_3"]


def species_quotas(p):
    """(concentrations[0..S], counts[S]) as initial_setup derives them (initialise.F90:468-506): species_numbers win when
    they sum to the number of sites (concentrations recomputed in single precision); otherwise nint(real(N) * c) with the
    rounding error spread one atom at a time over the species that are present."""
    S, N = p["n_species"], n_atoms(p)
    conc = list(p.get("species_concentrations", [0.0] * (S + 1)))
    numbers = p.get("species_numbers")
    if numbers is not None and sum(numbers) != N:
        # the reference falls through to the concentration branch with all-zero concentrations and never leaves its
        # redistribution loop (src/initialise.F90:468-506): refuse instead
        raise BrawlCudaError("species_numbers sum to %d, the lattice has %d sites" % (sum(numbers), N))
    if numbers is not None:
        count = list(numbers)
        conc = [0.0] + [float(np.float32(c) / np.float32(N)) for c in count]
    else:
        nint = lambda x: int(math.floor(x + 0.5)) if x >= 0 else -int(math.floor(-x + 0.5))
        count = [nint(float(np.float32(N)) * conc[i + 1]) for i in range(S)]
        chk = sum(count) - N
        inc = -1 if chk > 0 else 1
        if chk != 0 and not any(count):
            raise BrawlCudaError("species concentrations are all zero")
        while chk != 0:
            for j in range(S):
                if count[j] == 0 or chk == 0:
                    continue
                count[j] += inc
                chk += inc
    if sum(count) != N:
        raise BrawlCudaError("species counts do not sum to the number of lattice sites")
    return conc, count


def netcdf_setup(p):
    """The attribute fields the NetCDF writers need (brawl_b200/netcdf3.py)."""
    conc, _ = species_quotas(p)
    return dict(n_basis=p["n_basis"], n_1=p["n_1"], n_2=p["n_2"], n_3=p["n_3"], n_species=p["n_species"], lattice=p["lattice"],
                interaction_file=p["interaction_file"], species_concentrations=conc, wc_range=p["wc_range"])


def wang_landau_from_files(directory, walkers=8, **kw):
    """wl_main's set-up from the reference's own files in `directory`: brawl.inp, wl_input.inp and the interaction file."""
    from . import wang_landau as wl
    p = read_control_file(os.path.join(directory, "brawl.inp"))
    wp = wl.WLParams.from_file(os.path.join(directory, "wl_input.inp"))
    V = read_exchange(os.path.join(directory, p["interaction_file"]), p["n_species"], p["interaction_range"])
    _, counts = species_quotas(p)
    return wl.WangLandau(p["lattice"], p["n_1"], p["n_2"], p["n_3"], p["n_species"], p["interaction_range"], V, counts, wp,
                         walkers=walkers, **kw), p


def read_metropolis_file(path):
    """metropolis.inp -> dict (read_metropolis_file, src/io.f90:500-690): defaults, the n_sample_steps_* fall-backs
    (including the reference's asro/alro typo, io.f90:654-662) and `burn_in` implying `burn_in_start` (:681-684)."""
    if not os.path.exists(path):
        raise BrawlCudaError("Could not find Metropolis control file: " + path)
    m = dict(n_burn_in_steps=0, burn_in_start=False, burn_in=False, calculate_energies=True, calculate_asro=True,
             calculate_alro=False, n_sample_steps_asro=0, n_sample_steps_alro=0, n_sample_steps_trajectory=0,
             write_trajectory_xyz=False, write_trajectory_energy=False, write_trajectory_asro=False,
             write_initial_config_xyz=False, write_initial_config_nc=False, write_final_config_xyz=False,
             write_final_config_nc=False, read_start_config_nc=False, start_config_file="", T_steps=1, delta_T=1.0,
             nbr_swap=False)
    ints = ("n_mc_steps", "n_burn_in_steps", "n_sample_steps", "n_sample_steps_asro", "n_sample_steps_alro",
            "n_sample_steps_trajectory", "T_steps")
    for line in open(path):
        if line.lstrip().startswith("#") or "=" not in line:
            continue
        k, v = line.split("=", 1)
        k = k.rstrip(" \t")
        if k in ints:
            m[k] = int(v.split("#")[0])
        elif k in ("T", "delta_T"):
            m[k] = float(v.split("#")[0])
        elif k in ("mode", "start_config_file"):
            m[k] = _string(v)
        elif k in m and isinstance(m[k], bool):
            m[k] = _logical(v)
    for k in ("mode", "n_mc_steps", "n_sample_steps", "T"):
        if k not in m:
            raise BrawlCudaError("Missing '%s' in Metropolis input file" % k)
    if m["n_sample_steps_alro"] == 0:
        m["n_sample_steps_asro"] = m["n_sample_steps"]
        m["n_sample_steps_alro"] = m["n_sample_steps"]
    if m["n_sample_steps_trajectory"] == 0:
        m["n_sample_steps_trajectory"] = m["n_sample_steps"]
    if m["burn_in"]:
        m["burn_in_start"] = True
    return m


def replica_annealing_from_files(directory, n_replicas, **kw):
    """metropolis_simulated_annealing's set-up from brawl.inp + metropolis.inp + the interaction file, with
    `n_replicas` chains per GPU in place of one chain per MPI rank."""
    from . import replica_annealing as ra
    p = read_control_file(os.path.join(directory, "brawl.inp"))
    m = read_metropolis_file(os.path.join(directory, "metropolis.inp"))
    V = read_exchange(os.path.join(directory, p["interaction_file"]), p["n_species"], p["interaction_range"])
    _, counts = species_quotas(p)
    drv = ra.ReplicaAnnealing(p["lattice"], p["n_1"], p["n_2"], p["n_3"], p["n_species"], p["interaction_range"], V, counts,
                              n_replicas=n_replicas, T=m["T"], T_steps=m["T_steps"], delta_T=m["delta_T"],
                              n_mc_steps=m["n_mc_steps"], n_sample_steps=m["n_sample_steps"],
                              n_burn_in_steps=m["n_burn_in_steps"], burn_in_start=m["burn_in_start"], burn_in=m["burn_in"],
                              n_sample_steps_asro=m["n_sample_steps_asro"],
                              wc_range=p["wc_range"] if m["calculate_asro"] else 0, nbr_swap=m["nbr_swap"], **kw)
    return drv, p, m


def nested_sampling_from_files(directory, n_runs=1, **kw):
    """nested_sampling_main's set-up from brawl.inp + ns_input.inp + the interaction file."""
    from . import nested_sampling as ns
    p = read_control_file(os.path.join(directory, "brawl.inp"))
    sp = ns.NSParams.from_file(os.path.join(directory, "ns_input.inp"))
    V = read_exchange(os.path.join(directory, p["interaction_file"]), p["n_species"], p["interaction_range"])
    _, counts = species_quotas(p)
    return ns.NestedSampling(p["lattice"], p["n_1"], p["n_2"], p["n_3"], p["n_species"], p["interaction_range"], V, counts, sp,
                             n_runs=n_runs, **kw), p
