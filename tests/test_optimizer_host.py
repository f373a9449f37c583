"""Host logic of the Optimizer on CPU: which engine update each ask chooses (fit / append / rebase / reuse), with the engine
replaced by the oracle-backed stand-in from test_spmd_gloo (tests may use oracle/; the product has no CPU engine)."""
import numpy as np

from kubeflow_b200.optimizer import Optimizer
from kubeflow_b200.space import Categorical, Integer, Real
from tests.test_spmd_gloo import OracleEngine


def _opt(**kw):
    dims = [Real(0.0, 1.0, name="a"), Integer(1, 9, name="b"), Categorical(["x", "y", "z"], name="c"), Real(-2.0, 2.0, name="d")]
    eng = OracleEngine()
    return Optimizer(dims, n_initial_points=3, random_state=7, n_points=500, candidate_backend="numpy", engine=eng, **kw), eng


def _hist(n, seed=0):
    rng = np.random.default_rng(seed)
    X = [[float(rng.random()), int(rng.integers(1, 10)), str(rng.choice(["x", "y", "z"])), float(rng.uniform(-2, 2))] for _ in range(n)]
    y = [float(np.sin(3 * r[0]) + 0.1 * r[1] + (r[2] == "y") + r[3] ** 2) for r in X]
    return X, y


def test_engine_updates_follow_the_history():
    opt, eng = _opt()
    X, y = _hist(20)
    opt.tell(X, y)
    pts = opt.ask(n_points=3)
    assert eng.log == ["fit", "append", "append"] and opt.last_fit == "append"       # two constant lies appended
    for p in pts:
        assert 0 <= p[0] <= 1 and 1 <= p[1] <= 9 and p[2] in ("x", "y", "z") and -2 <= p[3] <= 2
    eng.log.clear()
    opt.tell([pts[0]], [0.5])            # the first lie became an observation at the same x: only the target changes
    opt.ask()
    assert eng.log == ["rebase"] and opt.last_fit == "rebase" and len(eng.y) == 21 and eng.y[-1] == 0.5
    eng.log.clear()
    opt.tell([pts[2]], [0.7])            # a new row after the common prefix
    opt.ask()
    assert eng.log == ["append"] and len(eng.y) == 22     # the engine held exactly the 21 rows: nothing to drop
    eng.log.clear()
    opt.ask()
    assert eng.log == [] and opt.last_fit == "reuse"
    opt.xi = 0.05                        # a different acquisition setting is a different engine state
    opt.ask()
    assert opt.last_fit == "fit"
    eng.log.clear()
    opt.Xi[2] = [0.5, 5, "x", 0.0]
    opt._Xt[2] = opt.space.transform([opt.Xi[2]])[0]     # an edited early row: no usable prefix
    opt.ask()
    assert eng.log == ["fit"]


def test_incremental_and_refit_paths_suggest_the_same_points():
    X, y = _hist(30, seed=3)
    outs = []
    for inc in (True, False):
        opt, eng = _opt(incremental=inc)
        opt.tell(X, y)
        a = opt.ask(n_points=4)
        opt.tell(a[:2], [0.1, 0.2])
        b = opt.ask(n_points=2)
        outs.append((a, b, list(eng.log)))
    assert outs[0][0] == outs[1][0] and outs[0][1] == outs[1][1]
    assert "append" in outs[0][2] and set(outs[1][2]) == {"fit"}


def test_no_room_in_the_pitch_means_a_refit():
    opt, eng = _opt()
    X, y = _hist(64, seed=5)             # 64 rows fill the 64-row pitch: room() == 0
    opt.tell(X, y)
    opt.ask(n_points=2)
    assert eng.log == ["fit", "fit"]     # the lie cannot be appended
    X2, y2 = _hist(1, seed=6)
    opt.tell(X2, y2)                     # 65 real rows: prefix of 64, one new row, room again after the refit
    eng.log.clear()
    opt.ask(n_points=2)
    assert eng.log[0] in ("fit", "rebase") and eng.log[-1] == "append"


def test_xt_passed_to_tell_must_be_what_transform_gives():
    opt, _ = _opt()
    X, y = _hist(5)
    xt = opt.space.transform(X)
    opt.tell(X, y, xt=xt)
    opt2, _ = _opt()
    opt2.tell(X, y)
    np.testing.assert_array_equal(opt._Xt, opt2._Xt)
    assert opt.Xi == opt2.Xi and opt.yi == opt2.yi
