"""The reference's text outputs (src/metropolis_output.f90) for the Python drivers: same file names' contents, same
Fortran edit descriptors.  Checked against the reference's golden files in tests/test_text_io.py."""
import os


def energy_trajectory_writer(filename, step, energy):
    """energy_trajectory_writer (metropolis_output.f90:34-53): appends '(I13,x,f20.10,x)' under ' # step_number E'."""
    new = not os.path.exists(filename)
    with open(filename, "a") as fh:
        if new:
            fh.write(" # step_number E\n")
        fh.write("%13d %20.10f\n" % (step, energy))


def asro_trajectory_writer(filename, step, asro):
    """asro_trajectory_writer (metropolis_output.f90:66-98): step then every r_densities entry as (f8.5,x), the
    array in Fortran element order (i fastest) = C order of rho[l][j][i]."""
    new = not os.path.exists(filename)
    with open(filename, "a") as fh:
        if new:
            fh.write(" # step_number ASRO\n")
        fh.write("%13d " % step + "".join("%8.5f " % v for v in asro) + "\n")


def diagnostics_writer(filename, temperature, energies, C, acceptance):
    """diagnostics_writer (metropolis_output.f90:113-135): '(F8.1,2X,F24.15,2X,F24.15,2X,F6.4)' per temperature."""
    with open(filename, "w") as fh:
        fh.write(" # T E C acceptance_rate\n")
        for T, E, c, a in zip(temperature, energies, C, acceptance):
            fh.write("%8.1f  %24.15f  %24.15f  %6.4f\n" % (T, E, c, a))
