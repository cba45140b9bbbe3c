"""Text writers of the Python drivers against the reference's golden trajectory / diagnostics files: every line of the
golden is parsed and written back; the files must be identical."""
import numpy as np


def test_trajectory_and_diagnostics_writers_reproduce_goldens(golden, tmp_path):
    from brawl_b200 import text_io
    for case in ("t01_", "t02_r0_", "t02_r3_"):
        ref = str(golden[case + "energy_txt"])
        p = str(tmp_path / (case + "e.dat"))
        for line in ref.split("\n")[1:]:
            if line.strip():
                text_io.energy_trajectory_writer(p, int(line.split()[0]), float(line.split()[1]))
        assert open(p).read() == ref
        ref = str(golden[case + "asro_txt"])
        p = str(tmp_path / (case + "a.dat"))
        for line in ref.split("\n")[1:]:
            if line.strip():
                text_io.asro_trajectory_writer(p, int(line.split()[0]), [float(x) for x in line.split()[1:]])
        assert open(p).read() == ref
    for key in ("t02_r0_diag_txt", "t02_av_diag_txt"):
        ref = str(golden[key])
        T, E, C, a = (float(x) for x in ref.split("\n")[1].split())
        p = str(tmp_path / (key + ".dat"))
        text_io.diagnostics_writer(p, [T], [E], [C], [a])
        assert open(p).read() == ref
