"""The Python NetCDF-3 classic writers (brawl_b200/netcdf3.py) against the reference's own golden files, byte for byte."""
import numpy as np
from scipy.io import netcdf_file


def _setup_from(f):
    return dict(n_1=f.N_1, n_2=f.N_2, n_3=f.N_3, n_species=getattr(f, "Number of Species"),
                lattice=getattr(f, "Lattice Type").decode(), interaction_file=getattr(f, "Interaction file").decode(),
                species_concentrations=np.array(f.Concentrations), wc_range=getattr(f, "Warren-Cowley Range"))


def test_radial_density_writer_reproduces_golden_files(golden, tmp_path):
    from brawl_b200 import netcdf3
    for key in ("raw_t02_r0_rho_nc", "raw_t02_av_rho_nc"):
        raw = golden[key].tobytes()
        src = str(tmp_path / "ref.nc")
        open(src, "wb").write(raw)
        f = netcdf_file(src, "r", mmap=False)
        v = {k: np.array(f.variables[k + " data"].data) for k in ("rho", "r", "T", "U")}
        out = str(tmp_path / "mine.nc")
        netcdf3.ncdf_radial_density_writer(out, v["rho"], v["r"], v["T"], v["U"], _setup_from(f))
        assert open(out, "rb").read() == raw, key


def test_writer_1d_and_across_energy(golden, tmp_path):
    from brawl_b200 import netcdf3
    p = str(tmp_path / "a.nc")
    netcdf3.ncdf_writer_1d(p, golden["t04_wl_dos"])
    assert open(p, "rb").read() == golden["raw_t04_wl_dos_nc"].tobytes()
    # rho(E): no golden in the reference; same header grammar as the pinned rho(T) writer, read back with scipy
    rng = np.random.default_rng(0)
    rho = rng.random((7, 3, 4, 4))
    setup = dict(n_1=4, n_2=4, n_3=4, n_species=4, lattice="bcc", interaction_file="bcc_epi.vij   ",
                 species_concentrations=[0.0, 0.25, 0.25, 0.25, 0.25], wc_range=3)
    netcdf3.ncdf_radial_density_writer_across_energy(p, rho, [0.0, 0.866, 1.0], np.linspace(-1, 0, 7), setup)
    f = netcdf_file(p, "r", mmap=False)
    assert list(f.dimensions.items()) == [("i", 4), ("j", 4), ("r", 3), ("U", 7), ("r_i", 3), ("U_i", 7)]
    assert f.variables["rho data"].dimensions == ("U", "r", "j", "i") and np.array_equal(f.variables["rho data"].data, rho)
    assert np.array_equal(f.variables["U data"].data, np.linspace(-1, 0, 7)) and getattr(f, "Interaction file") == b"bcc_epi.vij"
    assert getattr(f, "Warren-Cowley Range") == 3 and np.array_equal(f.Concentrations, [0.0, 0.25, 0.25, 0.25, 0.25])


def test_grid_state_writer_and_reader_against_golden(golden, tmp_path):
    """The reference's golden proc_0000_initial_config_at_0300.0.nc (case 02): read by the restart reader, rewritten
    byte for byte by the writer."""
    from brawl_b200 import netcdf3
    raw = golden["raw_t02_r0_initial_nc"].tobytes()
    src = str(tmp_path / "ref.nc")
    open(src, "wb").write(raw)
    cfg = netcdf3.ncdf_config_reader(src)
    assert cfg.dtype == np.int8 and np.array_equal(cfg, golden["t02_r0_initial"])
    f = netcdf_file(src, "r", mmap=False)
    setup = dict(n_basis=f.N_basis, n_1=f.N_1, n_2=f.N_2, n_3=f.N_3, n_species=getattr(f, "Number of Species"),
                 lattice=getattr(f, "Lattice Type").decode(), species_concentrations=np.array(f.Concentrations))
    out = str(tmp_path / "mine.nc")
    netcdf3.ncdf_grid_state_writer(out, cfg, setup)
    assert open(out, "rb").read() == raw
    import pytest
    open(out, "wb").write(b"HDF5....")
    with pytest.raises(ValueError):
        netcdf3.ncdf_config_reader(out)
