"""NetCDF-3 classic (CDF-1) writers for the Python drivers' outputs -- the on-disk contract of the reference's
src/netcdf_io.f90.  Pure host code, no netCDF library: fixed-size variables only, big-endian, header as
netCDF-C lays it out, so files are byte-identical to the reference's for the same data (tests compare against the
reference's own golden files).  The C++ host side (brawl_b200/host) has its own copy of the same format."""
import struct

import numpy as np

NC_CHAR, NC_SHORT, NC_INT, NC_DOUBLE = 2, 3, 4, 6
_NC_DIMENSION, _NC_VARIABLE, _NC_ATTRIBUTE = 10, 11, 12


def _pad(b):
    return b + b"\0" * (-len(b) % 4)


def _name(t):
    b = t.encode()
    return struct.pack(">I", len(b)) + _pad(b)


def _att(name, value):
    if isinstance(value, str):
        b = value.encode()
        return _name(name) + struct.pack(">II", NC_CHAR, len(b)) + _pad(b)
    a = np.atleast_1d(np.asarray(value))
    if a.dtype.kind in "iu":
        return _name(name) + struct.pack(">II", NC_INT, a.size) + a.astype(">i4").tobytes()
    return _name(name) + struct.pack(">II", NC_DOUBLE, a.size) + a.astype(">f8").tobytes()


def write_classic(filename, dims, gatts, variables):
    """dims: [(name, size)]; gatts: [(name, str | ints | floats)]; variables: [(name, dimids (file order, slowest
    first), NC_DOUBLE | NC_SHORT, data)].  Data are written in C order of the given arrays."""
    payload = []
    for _, _, typ, data in variables:
        a = np.ascontiguousarray(data)
        payload.append(_pad(a.astype(">f8" if typ == NC_DOUBLE else ">i2").tobytes()))

    def header(begins):
        h = b"CDF\x01" + struct.pack(">I", 0)
        h += struct.pack(">II", _NC_DIMENSION, len(dims)) if dims else struct.pack(">II", 0, 0)
        for n, size in dims:
            h += _name(n) + struct.pack(">I", size)
        h += struct.pack(">II", _NC_ATTRIBUTE, len(gatts)) if gatts else struct.pack(">II", 0, 0)
        for n, v in gatts:
            h += _att(n, v)
        h += struct.pack(">II", _NC_VARIABLE, len(variables)) if variables else struct.pack(">II", 0, 0)
        for (n, dimids, typ, _), pl, beg in zip(variables, payload, begins):
            h += _name(n) + struct.pack(">I", len(dimids)) + b"".join(struct.pack(">I", d) for d in dimids)
            h += struct.pack(">II", 0, 0) + struct.pack(">III", typ, len(pl), beg)
        return h

    hlen = len(header([0] * len(variables)))
    begins, off = [], hlen
    for pl in payload:
        begins.append(off)
        off += len(pl)
    with open(filename, "wb") as fh:
        fh.write(header(begins) + b"".join(payload))


def _setup_atts(setup):
    """The global attributes every radial-density file carries (netcdf_io.f90:181-198, 289-306)."""
    return [("N_1", int(setup["n_1"])), ("N_2", int(setup["n_2"])), ("N_3", int(setup["n_3"])),
            ("Number of Species", int(setup["n_species"])), ("Lattice Type", str(setup["lattice"]).rstrip()),
            ("Interaction file", str(setup["interaction_file"]).rstrip()),
            ("Concentrations", np.asarray(setup["species_concentrations"], dtype=np.float64)),
            ("Warren-Cowley Range", int(setup["wc_range"]))]


def ncdf_writer_1d(filename, grid_data):
    """ncdf_writer_1d (netcdf_io.f90:731-806): dimension "x", variable "grid data"."""
    data = np.asarray(grid_data, dtype=np.float64).ravel()
    write_classic(filename, [("x", data.size)], [], [("grid data", [0], NC_DOUBLE, data)])


def ncdf_radial_density_writer(filename, rho, r, T, U, setup):
    """ncdf_radial_density_writer (netcdf_io.f90:150-244): rho[T][r][j][i] (disk order of the Fortran rho(i,j,r,T))."""
    rho = np.asarray(rho, dtype=np.float64)
    nT, nr, S, _ = rho.shape
    write_classic(filename, [("i", S), ("j", S), ("r", nr), ("T", nT), ("r_i", len(r)), ("T_i", len(T)), ("U_i", len(U))],
                  _setup_atts(setup),
                  [("rho data", [3, 2, 1, 0], NC_DOUBLE, rho), ("r data", [4], NC_DOUBLE, r), ("T data", [5], NC_DOUBLE, T),
                   ("U data", [6], NC_DOUBLE, U)])


def ncdf_radial_density_writer_across_energy(filename, rho, r, U, setup):
    """ncdf_radial_density_writer_across_energy (netcdf_io.f90:260-346): rho[U][r][j][i] per energy bin (Wang-Landau
    rho(E), wang-landau.F90:374), r = shell radii, U = bin centre energies."""
    rho = np.asarray(rho, dtype=np.float64)
    nU, nr, S, _ = rho.shape
    write_classic(filename, [("i", S), ("j", S), ("r", nr), ("U", nU), ("r_i", len(r)), ("U_i", len(U))], _setup_atts(setup),
                  [("rho data", [3, 2, 1, 0], NC_DOUBLE, rho), ("r data", [4], NC_DOUBLE, r), ("U data", [5], NC_DOUBLE, U)])


def ncdf_grid_state_writer(filename, config, setup):
    """ncdf_grid_state_writer (netcdf_io.f90:495-583): one configuration, int8 grid [2n3][2n2][2n1] (the layout of
    Device.get_config) stored as NF90_SHORT "configuration"(z, y, x, b) -- the restart file ncdf_config_reader reads."""
    g = np.asarray(config)
    gz, gy, gx = g.shape
    atts = [("N_basis", int(setup.get("n_basis", 1))), ("N_1", int(setup["n_1"])), ("N_2", int(setup["n_2"])), ("N_3", int(setup["n_3"])),
            ("Number of Species", int(setup["n_species"])), ("Lattice Type", str(setup["lattice"]).rstrip()),
            ("Concentrations", np.asarray(setup["species_concentrations"], dtype=np.float64))]
    write_classic(filename, [("b", int(setup.get("n_basis", 1))), ("x", gx), ("y", gy), ("z", gz)], atts,
                  [("configuration", [3, 2, 1, 0], NC_SHORT, g.astype(np.int16))])


def ncdf_config_reader(filename):
    """ncdf_config_reader (netcdf_io.f90:1368-1429): the "configuration" variable of a classic file written by
    ncdf_grid_state_writer, returned as int8 [2n3][2n2][2n1] (ready for Device.set_config)."""
    d = open(filename, "rb").read()
    if d[:4] != b"CDF\x01":
        raise ValueError("%s is not a NetCDF classic file" % filename)
    rd = lambda o: struct.unpack(">I", d[o:o + 4])[0]
    o = 8
    dims = []
    if rd(o) == _NC_DIMENSION:
        n = rd(o + 4); o += 8
        for _ in range(n):
            ln = rd(o); o += 4 + ln + (-ln % 4)
            dims.append(rd(o)); o += 4
    else:
        o += 8
    if rd(o) == _NC_ATTRIBUTE:
        n = rd(o + 4); o += 8
        for _ in range(n):
            ln = rd(o); o += 4 + ln + (-ln % 4)
            typ, ne = rd(o), rd(o + 4); o += 8
            nb = ne * {NC_CHAR: 1, NC_SHORT: 2, NC_INT: 4, NC_DOUBLE: 8, 1: 1, 5: 4}[typ]
            o += nb + (-nb % 4)
    else:
        o += 8
    if rd(o) != _NC_VARIABLE:
        raise ValueError("%s holds no variables" % filename)
    n = rd(o + 4); o += 8
    for _ in range(n):
        ln = rd(o); name = d[o + 4:o + 4 + ln].decode(); o += 4 + ln + (-ln % 4)
        nd = rd(o); dimids = [rd(o + 4 + 4 * k) for k in range(nd)]; o += 4 + 4 * nd
        if rd(o) != 0 or rd(o + 4) != 0:
            raise ValueError("variable attributes are not expected in a configuration file")
        o += 8
        typ, _, begin = rd(o), rd(o + 4), rd(o + 8); o += 12
        if name == "configuration":
            if typ != NC_SHORT:
                raise ValueError("configuration is not NF90_SHORT")
            shape = [dims[k] for k in dimids]
            a = np.frombuffer(d, dtype=">i2", count=int(np.prod(shape)), offset=begin).reshape(shape)
            return a[..., 0].astype(np.int8)
    raise ValueError("%s holds no 'configuration' variable" % filename)
