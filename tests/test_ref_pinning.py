"""CPU suite, part 5: the oracle restatement (oracle/jackal_oracle.hpp) pinned against THE REFERENCE'S OWN CODE.

`oracle/_ref/libjackal_ref.so` is built from the reference's first-party sources where they lie under /root/reference
(HybridDynamics, TireNetwork, BekkerDynamics, BekkerTireModel, utils, the RobCoGen-generated forward_dynamics / transforms /
inertia_properties / miscellaneous, VehicleSystem, BekkerSystem, TestTerrainMaps, Trainer — recipe: oracle/Makefile `ref`)
against stand-in headers for its three absent third-party libraries (oracle/refshim/: Eigen dense algebra, CppAD AD<double> +
ADFun::Reverse, iit-rbd spatial helpers).  Every function of the hot path is compared here: the restatement must reproduce
the reference's own arithmetic to rounding level, including `Trainer::trainTrajectory`'s tape gradient.

The library is prebuilt in the build container and travels with the snapshot; where neither it nor /root/reference exists
the tests skip."""
import os

import numpy as np
import pytest

from oracle import pyoracle as po
from oracle import pyref as pr

pytestmark = pytest.mark.skipif(not (pr.available() or os.path.isdir(os.path.join(pr.REF, "src"))),
                                reason="oracle/_ref not built and /root/reference not mounted")

RNG = np.random.default_rng(20231115)
B = np.array([[0.0472, 0.0438], [0.0036, -0.0035], [-0.1744, 0.1748]])


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b))))


def random_state(z=0.0590):
    """a plausible driving state: small tilt, wheels in the soil, moderate velocities"""
    q = np.array([RNG.normal(0, 0.02), RNG.normal(0, 0.02), RNG.normal(0, 0.5), 1.0])
    q /= np.linalg.norm(q)
    X = np.zeros(21)
    X[0:4] = q
    X[4:7] = (RNG.uniform(1.5, 4.0), RNG.uniform(-1, 1), z + RNG.normal(0, 0.002))
    X[11:14] = RNG.normal(0, 0.3, 3)
    X[14:17] = (RNG.uniform(-0.2, 0.8), RNG.normal(0, 0.1), RNG.normal(0, 0.05))
    u = RNG.uniform(-1, 8, 2)
    X[17:21] = (u[0], u[1], u[0], u[1])
    return X, u


def window(T, x_shift=0.0, seed=0):
    """rows [T][10] = (time, vl, vr, x, y, z, yaw, wz, vx, vy): what Trainer::loadDataFile produces (ground truth by the
    reference's linear kinematic model, scripts/gen_fake_data.py)"""
    r = np.random.default_rng(seed)
    t = np.arange(T) * 0.01
    vl = r.uniform(1, 5) + np.sin(r.uniform(1, 3) * t + r.uniform(0, 6))
    vr = r.uniform(1, 5) + np.cos(r.uniform(1, 3) * t + r.uniform(0, 6))
    x, y, yaw = x_shift, 0.0, r.uniform(-3, 3)
    rows = np.zeros((T, 10))
    for i in range(T):
        v = B @ np.array([vl[i], vr[i]])
        rows[i] = (t[i], vl[i], vr[i], x, y, 0.0584427891, yaw, v[2], v[0], v[1])
        for _ in range(10):
            x += 1e-3 * (np.cos(yaw) * v[0] - np.sin(yaw) * v[1])
            y += 1e-3 * (np.sin(yaw) * v[0] + np.cos(yaw) * v[1])
            yaw += 1e-3 * v[2]
    return rows


def test_constants_and_default_parameters_are_the_references():
    assert pr.num_params(0) == 416 and pr.num_params(1) == 5
    assert np.array_equal(pr.default_params(0), po.nn_default_params())       # TireNetwork::load_model
    assert np.array_equal(pr.default_params(1), po.bekker_default_params())   # BekkerSystem::getDefaultParams
    np.testing.assert_allclose(pr.whole_body_com(), po.whole_body_com(), rtol=1e-15, atol=0)  # generated getWholeBodyCOM
    for terr in (0, 1, 2):
        for x in (-1.0, 1.9, 2.0, 2.3, 3.0, 3.59, 3.6, 3.7, 5.2):
            assert pr.altitude(terr, x, 0.4) == po.altitude(po.terrain(terr), x, 0.4)


def test_tire_network_forward_matches():
    P = pr.default_params(0)
    p2 = P + RNG.normal(0, 0.05, 416)
    worst = 0.0
    for params in (P, p2):
        for _ in range(50):
            in8 = np.zeros(8)
            in8[:4] = (RNG.normal(0.3, 0.5), RNG.normal(0, 0.2), RNG.uniform(-2, 10), RNG.uniform(-0.004, 0.02))
            worst = max(worst, rel(po.tire_nn_forward(params, in8), pr.tire_nn_forward(params, in8)))
    assert worst < 1e-13, worst


def test_generated_forward_dynamics_matches():
    """the RobCoGen-generated ABA (forward_dynamics.cpp + transforms.cpp + inertia_properties.cpp) with the iit-rbd helpers"""
    worst = 0.0
    for _ in range(40):
        v = RNG.normal(size=6)
        g = np.concatenate([np.zeros(3), RNG.normal(size=3) * 5])
        qd = RNG.normal(size=4) * 4
        fe = RNG.normal(size=(5, 6)) * 30
        a, qa = pr.fd(v, g, qd, fe)
        b, qb = po.fd(v, g, qd, fe)
        worst = max(worst, rel(b, a), rel(qb, qa))
    assert worst < 1e-12, worst


def test_bekker_tire_model_matches():
    soil = [29.76, 2083.0, 0.8, 0.05, 0.3927]
    worst = 0.0
    for _ in range(200):
        inp = [RNG.uniform(1e-4, 0.03), RNG.normal(0.3, 0.4), RNG.normal(0, 0.2), 0.098 * RNG.uniform(0.2, 9)]
        fr, wr = pr.bekker_forces(inp, soil)
        fo, wo = po.bekker_forces(inp, soil)
        worst = max(worst, rel(fo, fr), float(np.max(np.abs(wo - wr) / np.maximum(1.0, np.abs(wr)))))
    assert worst < 1e-11, worst


@pytest.mark.parametrize("tire", [0, 1])
def test_ode_and_integrate_match(tire):
    P = pr.default_params(tire)
    worst_ode = worst_int = 0.0
    for terr in (0, 1, 2):
        for _ in range(12):
            X, u = random_state(0.0590 if tire == 0 else 0.062)
            if terr == 1:
                X[6] += 0.1 * X[4]
            a = pr.ode(tire, P, terr, X, u)
            b = po.ode(tire, P, po.terrain(terr), X, u)
            worst_ode = max(worst_ode, rel(b, a))
            X23 = np.concatenate([X, u])
            worst_int = max(worst_int, rel(po.integrate(tire, P, po.terrain(terr), X23), pr.integrate(tire, P, terr, X23)))
    print(f"tire {tire}: ODE {worst_ode:.2e}, integrate {worst_int:.2e}")
    assert worst_ode < 1e-11 and worst_int < 1e-12


def test_contact_search_function_matches():
    """HybridDynamics::get_tire_cpts_sinkages_fancy (HybridDynamics.cpp:258-316), the reference's own function vs the restatement"""
    worst = 0.0
    for terr in (0, 1, 2):
        for _ in range(25):
            X, _ = random_state(0.055 + RNG.normal(0, 0.01))
            if terr == 1:
                X[6] += 0.1 * X[4]
            a, pa = pr.fancy_sinkages(terr, X)
            b, pb = po.fancy_sinkages(po.terrain(terr), X)
            worst = max(worst, float(np.abs(a - b).max()), float(np.abs(pa - pb).max()))
    assert worst < 1e-14, worst


def test_base_network_forward_matches():
    """BaseNetwork::forward (dead code in the reference's ODE, kept as API surface): the reference's own class vs the numpy restatement"""
    for _ in range(20):
        p = RNG.normal(0, 0.7, 131)
        x = RNG.normal(0, 2.0, 3)
        np.testing.assert_allclose(po.base_network_forward(p, x[:, None])[:, 0], pr.base_network_forward(p, x), rtol=1e-13, atol=1e-15)


def test_settle_and_adapters_match():
    for tire in (0, 1):
        P = pr.default_params(tire)
        s = pr.default_initial_state(tire, P, 0)
        so = po.settle(tire, P, po.terrain(0))
        assert rel(so[[0, 1, 2, 3, 6]], s[[0, 1, 2, 3, 6]]) < 1e-13
    row = window(3, seed=4)[1]
    np.testing.assert_allclose(po.initialize_state(row), pr.initialize_state(row), rtol=1e-15, atol=1e-18)
    X, u = random_state()
    gt = np.zeros(21)
    gt[3:6] = (0.3, X[4] + 0.2, X[5] - 0.1)
    assert po.loss(gt, np.concatenate([X, u])) == pytest.approx(pr.loss(gt, np.concatenate([X, u])), rel=1e-14)


def test_long_rollouts_match():
    """600-row windows (the reference's m_eval_steps): 5990 RK4 steps through the reference's integrate vs the restatement"""
    for tire, tol in ((0, 1e-11), (1, 1e-10)):
        P = pr.default_params(tire)
        rows = window(600, seed=11 + tire)
        if tire == 1:
            rows[:, 5] = 0.0604
        Lr, xr = pr.evaluate_trajectory(tire, P, 0, rows)
        Lo, xo = po.evaluate_trajectory(tire, P, po.terrain(0), rows)
        e = rel(xo, xr)
        print(f"tire {tire}: 600-row rollout, worst state difference {e:.2e}, loss {Lr:.6g} vs {Lo:.6g}")
        assert e < tol and Lo == pytest.approx(Lr, rel=1e-9)


@pytest.mark.parametrize("T,terr", [(8, 0), (8, 1), (8, 2), (40, 0), (40, 2)])
def test_train_trajectory_loss_and_gradient_match(T, terr):
    """Trainer::trainTrajectory — the reference's own loss, traj_len, tape structure and ADFun::Reverse(1) call — against the
    oracle's per-macro-step tape.  North star: gradients within 1e-7; achieved here: ~1e-15."""
    P = pr.default_params(0) + (0 if terr == 0 else np.concatenate([RNG.normal(0, 0.03, 104), np.zeros(312)]))
    rows = window(T, x_shift=2.2 if terr == 2 else 0.0, seed=T + terr)
    L, g, xl = pr.train_trajectory(P, terr, rows)
    Lo, go, xlo = po.train_trajectory(P, po.terrain(terr), rows)
    assert Lo == pytest.approx(L, rel=1e-12)
    assert rel(xlo, xl) < 1e-12
    scale = np.abs(g).max()
    assert np.max(np.abs(go - g)) / scale < 1e-12
    assert np.all(g[104:] == 0.0) and np.any(g[:104] != 0.0)  # only network 0 is evaluated (HybridDynamics.cpp:387)


def test_update_params_matches():
    """Trainer::updateParams (RMSProp on the float-truncated gradient).  The restatement is built with FMA contraction, the
    reference sources here without: one rounding of difference in 0.95 s + 0.05 g^2 is allowed, nothing more."""
    P = pr.default_params(0)
    g = RNG.normal(0, 3, 416)
    sq = np.abs(RNG.normal(1, 0.3, 416))
    p1, s1 = pr.update_params(P, sq, g)
    p2, s2 = po.rmsprop(P, sq, g)
    np.testing.assert_allclose(s2, s1, rtol=4e-16, atol=0)
    np.testing.assert_allclose(p2, p1, rtol=1e-15, atol=0)
    p1, s1 = pr.update_params(P, np.ones(416), g)   # the first step of a run (s = 1)
    p2, s2 = po.rmsprop(P, np.ones(416), g)
    np.testing.assert_allclose(p2, p1, rtol=1e-15, atol=0)
