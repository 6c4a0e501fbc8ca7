"""Known-answer tests for the restated SE3 (oracle/se3.cpp) on the element and tangent sets of
thirdparty/Sophus/sophus/test_se3.cpp, with the group-axiom checks of sophus/tests.hpp
(tolerance SophusConstants::epsilon() = 1e-10, scaled where the inputs are large)."""
import numpy as np
import oracle_py as O


def so3_exp(w):
    T = O.se3_exp(np.concatenate([np.zeros(3), w]))
    return T[:, :3]


def mk(R, t):
    return np.hstack([R, np.asarray(t, float)[:, None]])


def elements():
    e = [mk(so3_exp([0.2, 0.5, 0.0]), [0, 0, 0]), mk(so3_exp([0.2, 0.5, -1.0]), [10, 0, 0]), mk(so3_exp([0, 0, 0]), [0, 100, 5]),
         mk(so3_exp([0, 0, 0.00001]), [0, 0, 0]), mk(so3_exp([0, 0, 0.00001]), [0, -0.00000001, 0.0000000001]),
         mk(so3_exp([0, 0, 0.00001]), [0.01, 0, 0]), mk(so3_exp([np.pi, 0, 0]), [4, -5, 0])]
    a = O.se3_mul(O.se3_mul(mk(so3_exp([0.2, 0.5, 0.0]), [0, 0, 0]), mk(so3_exp([np.pi, 0, 0]), [0, 0, 0])), mk(so3_exp([-0.2, -0.5, -0.0]), [0, 0, 0]))
    b = O.se3_mul(O.se3_mul(mk(so3_exp([0.3, 0.5, 0.1]), [2, 0, -7]), mk(so3_exp([np.pi, 0, 0]), [0, 0, 0])), mk(so3_exp([-0.3, -0.5, -0.1]), [0, 6, 0]))
    return e + [a, b]


TANGENTS = [np.array(t, float) for t in ([0, 0, 0, 0, 0, 0], [1, 0, 0, 0, 0, 0], [0, 1, 0, 1, 0, 0], [0, -5, 10, 0, 0, 0],
                                         [-1, 1, 0, 0, 0, 1], [20, -1, 0, -1, 1, 0], [30, 5, -1, 20, -1, 0])]


def hat(x):
    v, w = x[:3], x[3:]
    M = np.zeros((4, 4))
    M[:3, :3] = [[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]
    M[:3, 3] = v
    return M


def vee(M):
    return np.array([M[0, 3], M[1, 3], M[2, 3], M[2, 1], M[0, 2], M[1, 0]])


def m44(T):
    return np.vstack([T, [0, 0, 0, 1]])


def test_rotation_is_orthonormal():
    for T in elements():
        R = T[:, :3]
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-12)
        assert abs(np.linalg.det(R) - 1) < 1e-12


def test_exp_log_roundtrip_on_group():  # tests.hpp expLogTest: exp(log(G)) == G
    for T in elements():
        T2 = O.se3_exp(O.se3_log(T))
        assert np.abs(T2 - T).max() < 1e-9 * max(1.0, np.abs(T).max())


def test_log_exp_roundtrip_on_tangent():  # tests.hpp expMapTest: log(exp(x)) == x (for |omega| < pi)
    for x in TANGENTS:
        if np.linalg.norm(x[3:]) >= np.pi:
            continue
        x2 = O.se3_log(O.se3_exp(x))
        assert np.abs(x2 - x).max() < 1e-9 * max(1.0, np.abs(x).max())


def test_exp_matches_matrix_exponential():
    from scipy.linalg import expm
    for x in TANGENTS:
        T = O.se3_exp(x)
        E = expm(hat(x))
        assert np.abs(m44(T) - E).max() < 1e-8 * max(1.0, np.abs(E).max())


def test_adjoint():  # tests.hpp adjointTest: hat(Ad_T x) == T hat(x) T^-1
    for T in elements():
        A = O.se3_adj(T)
        Ti = m44(O.se3_inv(T))
        for x in TANGENTS:
            lhs = A @ x
            rhs = vee(m44(T) @ hat(x) @ Ti)
            assert np.abs(lhs - rhs).max() < 1e-8 * max(1.0, np.abs(rhs).max())


def test_group_action_and_inverse():  # tests.hpp groupActionTest + inverse
    p = np.array([1.0, 2.0, 4.0])
    for T in elements():
        Ti = O.se3_inv(T)
        I = O.se3_mul(T, Ti)
        assert np.abs(I - np.eye(4)[:3]).max() < 1e-9 * max(1.0, np.abs(T).max())
        q = T[:, :3] @ p + T[:, 3]
        back = Ti[:, :3] @ q + Ti[:, 3]
        assert np.abs(back - p).max() < 1e-9 * max(1.0, np.abs(q).max())


def test_mul_matches_matrix_product():
    els = elements()
    for A in els:
        for B in els:
            Cm = O.se3_mul(A, B)
            ref = (m44(A) @ m44(B))[:3]
            assert np.abs(Cm - ref).max() < 1e-9 * max(1.0, np.abs(ref).max())


def test_ldlt_solve_spd():
    rng = np.random.default_rng(3)
    for n in (6, 8, 60):
        M = rng.normal(size=(n, n))
        A = M @ M.T + n * np.eye(n)
        b = rng.normal(size=n)
        x = O.ldlt_solve(A, b)
        assert np.allclose(A @ x, b, atol=1e-9)
