#!/usr/bin/env python3
"""Regenerate models/*.yaml from the model constants.

The reference ships four model files (models/model_*_params.yaml: `type`, `frequency`, flat square
`Q`, `R`, `P`) written by matlab/matlab2yaml.m with "%.20f".  The constants are
Q = Gamma diag(sigma_a^2) Gamma^T with Gamma = [dt^2/2 I; dt I; (I)], dt = 1/250, sigma_a = 1e-3 (linear) /
1e-5 (angular), R = diag(sigma_m^2) with sigma_m = 0.01 m / 0.1 rad, P = diag(sigma_p); "%.20f"
quantises the smallest entry (6.4e-21 -> 1e-20), which is reproduced by formatting the same way.
tests/test_models.py checks (where /root/reference is mounted) that the parsed values are
bit-identical to the reference files.
"""
import os
import numpy as np

DT = 1.0 / 250.0


def gamma_q(nb, axes_sigma):
    """nb kinematic blocks (2 = p,v ; 3 = p,v,a) over len(axes_sigma) axes."""
    na = len(axes_sigma)
    n = nb * na
    col = [0.5 * DT * DT, DT, 1.0][:nb] if nb == 3 else [0.5 * DT * DT, DT]
    Q = np.zeros((n, n))
    for ax, s in enumerate(axes_sigma):
        for a in range(nb):
            for b in range(nb):
                Q[a * na + ax, b * na + ax] = col[a] * col[b] * s * s
    return Q


def fmt(v):
    return "[" + ", ".join("%.20f" % x for x in v) + "]"


def write(path, typ, Q, R, P):
    with open(path, "w") as f:
        f.write("type: %s\n" % typ)
        f.write("frequency: %f\n" % (1.0 / DT))
        # matlab2yaml.m writes the matrix row by row
        f.write("Q: %s\n" % fmt(Q.reshape(-1)))
        f.write("R: %s\n" % fmt(R.reshape(-1)))
        f.write("P: %s\n" % fmt(P.reshape(-1)))


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    lin, ang = 1e-3, 1e-5
    R3 = np.diag([0.01 ** 2] * 3)
    R6 = np.diag([0.01 ** 2] * 3 + [0.1 ** 2] * 3)
    write(os.path.join(out_dir, "model_uniform_velocity_params.yaml"), "uniform_velocity",
          gamma_q(2, [lin] * 3), R3, np.diag([0.1] * 3 + [0.01] * 3))
    write(os.path.join(out_dir, "model_uniform_acceleration_params.yaml"), "uniform_acceleration",
          gamma_q(3, [lin] * 3), R3, np.diag([0.1] * 3 + [0.01] * 3 + [0.001] * 3))
    write(os.path.join(out_dir, "model_angular_velocities_params.yaml"), "angular_velocities",
          gamma_q(2, [lin] * 3 + [ang] * 3), R6, np.diag([0.1] * 3 + [0.01] * 9))
    write(os.path.join(out_dir, "model_angular_rates_params.yaml"), "angular_rates",
          gamma_q(3, [lin] * 3 + [ang] * 3), R6, np.diag([0.1] * 3 + [0.01] * 15))


if __name__ == "__main__":
    main(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "models"))
