"""TEST INFRASTRUCTURE -- the single-step vectors of the reference's own tests, run through the UNMODIFIED
reference: the analytic in-cell step (``tests/test_particle.py:75-209``: the two regression cases of issue 55,
under-flowing velocities and velocity differences, cells with closed walls and with impossible inflow) as
``_time2wall`` -> ``_which_early`` -> ``_increment`` compute it (``seaduck/utils.py:859,875,777`` as called by
``Particle.analytical_step``, ``lagrangian.py:561``), for both directions of time and with a stop in reach.

    python oracle/make_golden_known_answers.py        # writes tests/golden/analytic_step.npz

The reference's random cases draw from numpy's global generator; here every case seeds it first, so the vectors
are reproducible (the arrays are stored anyway).
"""

import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.refload import load_reference  # noqa: E402

sd = load_reference()
lag = importlib.import_module("seaduck.lagrangian")
utils = importlib.import_module("seaduck.utils")


def random_state(seed):
    """tests/test_particle.py:44 (random_p): u, du, rx, v, dv, ry, w, dw, rzl_lin"""
    np.random.seed(seed)
    vals = np.random.uniform(-0.5, 0.5, 9)
    u, du, rx, v, dv, ry, w, dw, rzl = vals
    rzl += 0.5
    return dict(u=np.array([u, v, w]) / 1e5, du=np.array([du, dv, dw]) / 1e5, x=np.array([rx, ry, rzl - 0.5]))


def from_walls(state, walls):
    """tests/test_particle.py:68 (u_du_from_uwall) for the three axes"""
    x = state["x"]
    u = np.array([walls[2 * i] * (0.5 - x[i]) + walls[2 * i + 1] * (x[i] + 0.5) for i in range(3)])
    du = np.array([walls[2 * i + 1] - walls[2 * i] for i in range(3)])
    return dict(u=u, du=du, x=x)


def cases():
    out = []
    out.append(("issue55", dict(u=np.array([4.695701621346472e-06, -1.4674619307268205e-05, 0.0]),
                                du=np.array([-5.773571643116384e-05, -2.8246484204894833e-06, -2.8246484204894833e-06]),
                                x=np.array([-0.009470161238869457, 0.08627583529220939, 0.3454546420042159]))))
    out.append(("corner", dict(u=np.array([0.00023098, -4.54496561e-06, 0.0]), du=np.array([-1.58748431e-05, 1.31782954e-05, 1.31782954e-05]),
                               x=np.array([-0.5, -0.5, 0.34545464]))))
    for seed in (43, 20, 628):
        s = random_state(seed)
        s["du"] = s["du"] / 1e5
        out.append((f"underflow_du_{seed}", s))
    for seed in (432, 320, 60288):
        s = random_state(seed)
        s["u"] = s["u"] / 1e5
        out.append((f"underflow_u_{seed}", s))
    for seed in range(1999, 2011):  # test_impossible_wall :134
        np.random.seed(seed + 100000)
        not_wall = np.random.randint(0, 2, 6)
        n_open = int(not_wall.sum())
        walls = np.zeros(6)
        if n_open:
            vels = np.random.uniform(-1e-5, 1e-5, n_open)
            vels[-1] = -np.sum(vels[:-1])
            walls[not_wall.astype(bool)] = vels
        walls *= np.array([1, -1, 1, -1, 1, -1])
        out.append((f"closed_walls_{seed}", from_walls(random_state(seed), walls)))
    for seed in range(2011, 2023):  # test_impossible_inflow :160
        np.random.seed(seed + 100000)
        not_wall = np.random.randint(0, 2, 6)
        n_open = int(not_wall.sum())
        walls = np.random.random(6)
        walls[~not_wall.astype(bool)] = 1
        if n_open:
            vels = np.random.uniform(-1e-5, 1e-5, n_open)
            vels[-1] = -np.sum(vels[:-1])
            walls[not_wall.astype(bool)] = vels
        walls *= np.array([1, -1, 1, -1, 1, -1])
        out.append((f"inflow_{seed}", from_walls(random_state(seed), walls)))
    for seed in range(2023, 2028):  # test_cross_cell_wall :187
        out.append((f"plain_{seed}", random_state(seed)))
    return out


def reference_step(state, tf):
    u_list = [np.array([v]) for v in state["u"]]
    du_list = [np.array([v]) for v in state["du"]]
    pos_list = [np.array([v]) for v in state["x"]]
    tfa = np.array([float(tf)])
    with np.errstate(all="ignore"):
        ts = lag._time2wall(pos_list, u_list, du_list, tfa)
        tend, t_event = lag._which_early(tfa, ts)
        newx, newu = [], []
        for i in range(3):
            move = utils._increment(t_event, u_list[i], du_list[i])
            newu.append(float((u_list[i] + du_list[i] * move)[0]))
            newx.append(float((move + pos_list[i])[0]))
    return int(tend[0]), float(t_event[0]), newx, newu


def main():
    names, x0, u, du, tf, tend, tev, newx, newu = [], [], [], [], [], [], [], [], []
    for name, st in cases():
        for t_stop in (1e80, -1e80, 2.0e3, -3.0e3):
            a, b, c, d = reference_step(st, t_stop)
            names.append(name)
            x0.append(st["x"]), u.append(st["u"]), du.append(st["du"]), tf.append(t_stop)
            tend.append(a), tev.append(b), newx.append(c), newu.append(d)
    path = os.path.join(ROOT, "tests", "golden", "analytic_step.npz")
    np.savez_compressed(path, name=np.array(names), x0=np.array(x0), u=np.array(u), du=np.array(du), tf=np.array(tf),
                        tend=np.array(tend, np.int32), t_event=np.array(tev), newx=np.array(newx), newu=np.array(newu))
    tend = np.array(tend)
    print(f"{len(names)} steps -> {os.path.getsize(path) / 1e3:.0f} kB; tend histogram {np.bincount(tend, minlength=7).tolist()}")


if __name__ == "__main__":
    main()
