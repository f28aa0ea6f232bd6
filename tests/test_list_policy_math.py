"""Host-side property checks of the neighbour-list validity rules the kernel applies (list_policy / build_list /
refine_list in cvo_rgbd_b200/csrc/cvo_lists.cuh), in f64 on random clouds and poses.

The lists replace the kd-tree radius search of the reference (src/cvo.cpp:106-125): they must hold a SUPERSET of
the pairs that can pass the strict gates at every pose they are used for, otherwise a nonzero of A would be lost."""
import numpy as np


def rand_rot(rng, angle):
    a = rng.normal(size=3)
    a /= np.linalg.norm(a)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * K @ K


def tf(R, T):
    """update_tf (src/cvo.cpp:83-87): y -> R^T y - R^T T."""
    return R.T, -R.T @ T


def own_radius2(t_c, t_lim, ell):
    """Squared ball radius of a pair with colour exponent t_c: a > sp_thres <=> d2 log2(e)/(2 l^2) + t_c < T."""
    return np.maximum(t_lim - t_c, 0.0) * 2 * ell * ell / np.log2(np.e)


def setup(rng, n=300):
    x = rng.uniform(-1, 1, size=(n, 3))
    y0 = x + rng.normal(size=(n, 3)) * 0.03
    t_c = rng.uniform(0, 0.4, size=(n, n))  # pose-independent colour exponents
    return x, y0, t_c


def test_displacement_bound_is_attained_at_a_bounding_box_corner():
    """disp = max_j |(M1 - M0) y_j + (t1 - t0)| is bounded by the same expression at the 8 box corners (convexity)."""
    rng = np.random.default_rng(1)
    for _ in range(50):
        _, y0, _ = setup(rng)
        M0, t0 = tf(rand_rot(rng, 0.05), rng.normal(size=3) * 0.05)
        M1, t1 = tf(rand_rot(rng, 0.07), rng.normal(size=3) * 0.05)
        lo, hi = y0.min(0), y0.max(0)
        corners = np.array([[(hi if (c >> k) & 1 else lo)[k] for k in range(3)] for c in range(8)])
        bound = np.linalg.norm(corners @ (M1 - M0).T + (t1 - t0), axis=1).max()
        true = np.linalg.norm(y0 @ (M1 - M0).T + (t1 - t0), axis=1).max()
        assert true <= bound * (1 + 1e-12)


def test_list_built_with_a_skin_covers_every_pair_that_passes_while_the_policy_holds():
    rng = np.random.default_rng(2)
    t_lim, skin = 0.32, 0.08
    used = 0
    for trial in range(30):
        x, y0, t_c = setup(rng)
        ell0 = rng.uniform(0.08, 0.15)
        r0 = np.sqrt(t_lim * 2 * ell0 * ell0 / np.log2(np.e))
        s = skin * r0
        R0, T0 = rand_rot(rng, 0.02), rng.normal(size=3) * 0.02
        M0, t0 = tf(R0, T0)
        d0 = np.linalg.norm(x[:, None, :] - (y0 @ M0.T + t0)[None, :, :], axis=2)
        listed = d0 < np.sqrt(own_radius2(t_c, t_lim, ell0)) + s
        for _ in range(10):  # later poses / length-scales: used only while max(0, r1 - r0) + disp <= s
            M1, t1 = tf(rand_rot(rng, rng.uniform(0, 0.6) * s) @ R0, T0 + rng.normal(size=3) * 0.25 * s)
            ell1 = ell0 * rng.uniform(0.6, 1.02)
            r1 = np.sqrt(t_lim * 2 * ell1 * ell1 / np.log2(np.e))
            disp = np.linalg.norm(y0 @ (M1 - M0).T + (t1 - t0), axis=1).max()
            if max(0.0, r1 - r0) + disp > s:
                continue
            used += 1
            d1 = np.linalg.norm(x[:, None, :] - (y0 @ M1.T + t1)[None, :, :], axis=2)
            passing = d1 * d1 < own_radius2(t_c, t_lim, ell1)  # includes the strict ell-ball (own radius <= r1)
            assert not np.any(passing & ~listed), trial
            assert passing.sum() > 0
    assert used >= 60  # (the motions above are sized so that the policy holds for a good share of the poses)


def test_narrowing_in_place_keeps_the_superset_property():
    """refine_list: with the ball shrunk (r1 <= r0) and s1 + disp <= s0, the pairs within r_e1 + s1 of pose 1 are all in
    the old list, so filtering the old list gives exactly the list a rebuild at pose 1 would give."""
    rng = np.random.default_rng(3)
    t_lim, skin = 0.32, 0.08
    used = 0
    for trial in range(30):
        x, y0, t_c = setup(rng)
        ell0 = rng.uniform(0.08, 0.15)
        r0 = np.sqrt(t_lim * 2 * ell0 * ell0 / np.log2(np.e))
        s0 = skin * r0
        R0, T0 = rand_rot(rng, 0.02), rng.normal(size=3) * 0.02
        M0, t0 = tf(R0, T0)
        y_at_0 = y0 @ M0.T + t0
        d0 = np.linalg.norm(x[:, None, :] - y_at_0[None, :, :], axis=2)
        old = d0 < np.sqrt(own_radius2(t_c, t_lim, ell0)) + s0
        ell1 = ell0 * rng.uniform(0.5, 0.7)
        r1 = np.sqrt(t_lim * 2 * ell1 * ell1 / np.log2(np.e))
        M1, t1 = tf(rand_rot(rng, rng.uniform(0, 0.3) * s0) @ R0, T0 + rng.normal(size=3) * 0.1 * s0)
        disp = np.linalg.norm(y0 @ (M1 - M0).T + (t1 - t0), axis=1).max()
        s1 = min(skin * r1, s0 - disp)
        if s1 < 0.25 * skin * r1:
            continue
        used += 1
        d1 = np.linalg.norm(x[:, None, :] - (y0 @ M1.T + t1)[None, :, :], axis=2)
        rebuilt = d1 < np.sqrt(own_radius2(t_c, t_lim, ell1)) + s1
        assert not np.any(rebuilt & ~old), trial          # nothing a rebuild would list is missing from the old list
        narrowed = old & rebuilt                           # what the filter keeps
        assert np.array_equal(narrowed, rebuilt) and rebuilt.sum() > 0
    assert used >= 15
