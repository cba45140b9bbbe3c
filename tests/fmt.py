"""Fortran edit-descriptor formatting used by the reference's text writers
(src/metropolis_output.f90:49, 84-94, 127-131); test helper."""


def energy_line(step, e):          # "(I13,x,f20.10,x)"
    return "%13d %20.10f" % (step, e)   # trailing x emits nothing


def energy_trajectory(energies, steps=None):
    steps = range(len(energies)) if steps is None else steps
    return " # step_number E\n" + "".join(energy_line(s, e) + "\n" for s, e in zip(steps, energies))


def asro_line(step, rho_lji):
    # I13,x then (f8.5,x) over asro(i,j,k) with i fastest == disk order [l][j][i] flattened; then x
    return "%13d " % step + "".join("%8.5f " % v for v in rho_lji.ravel())


def diagnostics(temps, energies, C, acc):   # '(F8.1,2X,F24.15,2X,F24.15,2X,F6.4)'
    return " # T E C acceptance_rate\n" + "".join(
        "%8.1f  %24.15f  %24.15f  %6.4f\n" % t for t in zip(temps, energies, C, acc))
