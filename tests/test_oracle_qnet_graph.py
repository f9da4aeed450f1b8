"""Pins oracle/qnet_oracle.py (what every Q-network GPU test compares against) to the reference's OWN TensorFlow graphs.

The reference ships no weights and TensorFlow 1.12 cannot be installed, so real TF outputs do not exist ("parity unpinned"
for TF's floating-point kernels, stated in DESIGN.md).  What does survive is the serialized GraphDef inside every
checkpoint's ``.meta`` file: the forward ops with their attributes, the loss, the complete gradient sub-graph that
``tf.gradients`` generated and the ``ApplyAdam`` ops with their constants.  ``tests/golden/ref_graph_*.json`` are those graphs
(parsed by ``oracle/tf_graph.py``, written by ``tests/golden/make_golden.py graphs``); ``oracle/tf_graph.GraphRunner`` executes
them op by op in NumPy float64.  The tests below

  * re-parse the ``.meta`` files when ``/root/reference`` is present and require the committed fixtures to be identical,
  * read the static facts off the graph (op sequence, strides, paddings, variable shapes in creation order, Adam constants,
    initialiser stddev) and compare them with ``qnet_oracle.layout`` / ``AdamTF1`` / the CUDA library's layout,
  * EXECUTE the reference graph -- Q-values, cost, every gradient tensor that reaches an ApplyAdam op, one optimizer step with
    the beta-power update -- on seeded inputs and compare with the oracle (float64 vs float64: 1e-9 relative).

Reference: BrainDQN.py:119-172 (vanilla, reduce_sum cost), BrainDQNNature.py:35-123 (eval_net / target_net, reduce_mean cost,
target_replace_op), BrainDoubleDQN.py (saves the very same graph as Nature: SURVEY quirk Q1).
"""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import qnet_oracle as qo  # noqa: E402
from oracle import tf_graph as tg  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
REF = os.environ.get("FLAPPY_REFERENCE", "/root/reference")
VAR_ORDER = ["w1", "b1", "w2", "b2", "w3", "b3", "wf1", "bf1", "wf2", "bf2"]        # TF variable creation order (BrainDQN.py:121-153)


@pytest.fixture(scope="module")
def graphs():
    return {"dqn": tg.load_json(os.path.join(GOLD, "ref_graph_dqn.json")),
            "dqn_nature": tg.load_json(os.path.join(GOLD, "ref_graph_dqn_nature.json"))}


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "train_history")), reason="reference checkout not present")
def test_fixtures_are_the_reference_meta_files(graphs):
    for key, g in graphs.items():
        fresh = tg.parse_meta(os.path.join(REF, g["source"]))
        assert fresh["nodes"] == g["nodes"] and fresh["tf_version"] == g["tf_version"] == "1.12.0", key
    # --model ddqn saved the same graph as dqnnature (its trainQNetwork is never called: SURVEY Q1), and every checkpoint of a
    # run holds the same graph
    dd = tg.parse_meta(os.path.join(REF, "train_history/double_dqn/bird-1500000.meta"))
    assert dd["nodes"] == graphs["dqn_nature"]["nodes"]
    other = tg.parse_meta(os.path.join(REF, "train_history/dqn_nature/bird-1300000.meta"))
    assert other["nodes"] == graphs["dqn_nature"]["nodes"]


@pytest.mark.parametrize("key,prefix", [("dqn", ""), ("dqn_nature", "eval_net/")])
def test_static_graph_facts_match_the_restatement(graphs, key, prefix):
    s = tg.summarize(graphs[key])
    assert s["tf_version"] == "1.12.0"
    # K1..K6 of SURVEY 2.2 in order: conv s4 + relu, ONE max-pool, conv s2 + relu, conv s1 + relu, reshape, fc + relu, fc
    core = [op for op in s["forward_ops"] if op != "Add" or True]
    assert [op for op in core if op in ("Conv2D", "MaxPool", "MatMul", "Reshape")] == ["Conv2D", "MaxPool", "Conv2D", "Conv2D", "Reshape", "MatMul", "MatMul"]
    assert core.count("Relu") == 4
    nets = 2 if key == "dqn_nature" else 1
    assert [c["strides"] for c in s["convs"]] == [[1, 4, 4, 1], [1, 2, 2, 1], [1, 1, 1, 1]] * nets
    assert all(c["padding"] == "SAME" and c["data_format"] == "NHWC" for c in s["convs"])
    assert all(p["ksize"] == [1, 2, 2, 1] and p["strides"] == [1, 2, 2, 1] and p["padding"] == "SAME" for p in s["pools"]) and len(s["pools"]) == nets
    # variables: creation order and shapes = the flat layout of the oracle and of the CUDA library (fb_qnet.cuh qnet_layout)
    L = qo.layout(512, False)
    model_vars = [v for v in s["variables"] if "Adam" not in v["name"] and "power" not in v["name"]]
    assert [v["name"] for v in model_vars[:10]] == [prefix + ("Variable" if i == 0 else f"Variable_{i}") for i in range(10)]
    assert [tuple(v["shape"]) for v in model_vars[:10]] == [tuple(L[n][1]) for n in VAR_ORDER]
    off = 0
    for n in VAR_ORDER:
        assert L[n][0] == off, n
        off += int(np.prod(L[n][1]))
    assert off == L["total"] == 898722
    if key == "dqn_nature":                                    # target net: same shapes, created after the whole eval net
        assert [tuple(v["shape"]) for v in model_vars[10:20]] == [tuple(L[n][1]) for n in VAR_ORDER]
        assert all(v["name"].startswith("target_net/") for v in model_vars[10:20])
    # every trainable variable has its m and v slot, and only the online net is optimised
    slots = [v["name"] for v in s["variables"] if v["name"].endswith("/Adam") or v["name"].endswith("/Adam_1")]
    assert len(slots) == 20 and all((prefix + "Variable") in n for n in slots)
    # tf.train.AdamOptimizer(1e-6): fp32 constants as stored in the graph == AdamTF1's
    a, ref = s["adam"], qo.AdamTF1(4)
    assert np.float32(a["lr"]) == ref.lr == np.float32(1e-6) and np.float32(a["beta1"]) == ref.b1 == np.float32(0.9)
    assert np.float32(a["beta2"]) == ref.b2 == np.float32(0.999) and np.float32(a["epsilon"]) == ref.eps == np.float32(1e-8)
    assert a["use_nesterov"] is False
    assert all(np.float32(v) == np.float32(0.01) for v in s["truncated_normal_stddev"])     # weight init stddev (BrainDQN.py:122)
    assert s["op_histogram"]["ApplyAdam"] == 10 and s["op_histogram"]["MaxPoolGrad"] == 1
    assert s["op_histogram"]["Conv2DBackpropFilter"] == 3 and s["op_histogram"]["Conv2DBackpropInput"] == 3


def _variables(prefix, flat):
    L = qo.layout(512, False)
    out = {}
    for i, n in enumerate(VAR_ORDER):
        o, shp = L[n]
        out[prefix + ("Variable" if i == 0 else f"Variable_{i}")] = flat[o:o + int(np.prod(shp))].astype(np.float64).reshape(shp)
    return out


def _minibatch(B, seed):
    rng = np.random.default_rng(seed)
    x = (rng.random((B, 5, 80, 80)) < 0.25).astype(np.uint8) * 255           # observations are exactly {0, 255}
    a = rng.integers(0, 2, B).astype(np.uint8)
    r = rng.choice(np.array([0.1, 3.0, -3.0], np.float32), B, p=[0.7, 0.15, 0.15])
    term = (r == -3.0).astype(np.uint8)
    return x, a, r, term


def _nhwc(frames):          # u8 [B,4,80,80] (channel = time, oldest first) -> the [B,80,80,4] array the reference feeds
    return np.transpose(frames, (0, 2, 3, 1)).astype(np.float64)


@pytest.mark.parametrize("key,variant,loss_sum,prefix,scope", [("dqn", 0, True, "", ""), ("dqn_nature", 1, False, "eval_net/", "target_net/")])
def test_executing_the_reference_graph_matches_the_oracle(graphs, key, variant, loss_sum, prefix, scope):
    """forward Q(s), Q_target(s'), cost, every gradient, one Adam step: the reference's GraphDef interpreted in NumPy float64
    against oracle/qnet_oracle.py (torch float64 autograd)"""
    g = graphs[key]
    B = 6
    p = qo.init_params(512, False, seed=11) * np.float32(3.0)
    t = qo.init_params(512, False, seed=12) * np.float32(3.0)
    x, a, r, term = _minibatch(B, 3)
    s, s2 = x[:, 0:4], x[:, 1:5]
    variables = _variables(prefix, p)
    if key == "dqn_nature":
        variables.update(_variables("target_net/", t))
    L = qo.layout(512, False)
    run = tg.GraphRunner(g, variables)
    q_node = prefix + "add_4"
    state_ph = prefix + "Placeholder"
    # ---- forward (BrainDQN.py:100 QValue.eval; BrainDQNNature.py:163 readout_t.eval)
    q_s, = run.run([q_node], {state_ph: _nhwc(s)})
    q_ref = qo.forward(torch.tensor(p.astype(np.float64)), s).numpy()
    assert np.abs(q_s - q_ref).max() <= 1e-9 * np.abs(q_ref).max()
    if key == "dqn_nature":
        q_next, = run.run(["target_net/add_4"], {"target_net/Placeholder": _nhwc(s2)})
        assert np.abs(q_next - qo.forward(torch.tensor(t.astype(np.float64)), s2).numpy()).max() <= 1e-9 * np.abs(q_next).max()
    # ---- the oracle's update, whose y (built in Python float64, fed as fp32) the graph receives as q_target
    loss_ref, g_ref, ae_ref, y, _ = qo.loss_and_grads(variant, p, t, s, s2, a, r, term, None, 0.99, loss_sum)
    onehot = np.eye(2)[a.astype(np.int64)]
    feeds = {state_ph: _nhwc(s), scope + "Placeholder_1": onehot, scope + "Placeholder_2": y.astype(np.float64)}
    cost_node = "Sum_1" if key == "dqn" else "target_net/Mean"
    cost, = run.run([cost_node], feeds)
    assert abs(float(cost) - loss_ref) <= 1e-9 * abs(loss_ref), (float(cost), loss_ref)
    # ---- gradients: the tensors the graph hands to its ApplyAdam ops
    by = {n["name"]: n for n in g["nodes"]}
    train_op = scope + "Adam"
    applies = [i[1:] for i in by[train_op]["inputs"] if i.startswith("^") and by[i[1:]]["op"] == "ApplyAdam"]
    assert len(applies) == 10
    run.feeds, run.memo = feeds, {}
    for i, name in enumerate(VAR_ORDER):
        nd = by[[ap for ap in applies if by[ap]["inputs"][0] == prefix + ("Variable" if i == 0 else f"Variable_{i}")][0]]
        grad = run.value(nd["inputs"][9])
        o, shp = L[name]
        ref = g_ref[o:o + int(np.prod(shp))].reshape(shp)
        err = np.linalg.norm(grad - ref) / np.linalg.norm(ref)
        assert err <= 1e-9, (name, err)
    # ---- one optimizer step (ApplyAdam x 10, then beta1_power / beta2_power <- power * beta), twice, against AdamTF1 run in float64
    slots = {}
    for ap in applies:
        _, m_name, v_name = [i.split(":")[0] for i in by[ap]["inputs"][:3]]
        shape = variables[by[ap]["inputs"][0]].shape
        slots[m_name] = np.zeros(shape); slots[v_name] = np.zeros(shape)
    run.vars.update(slots)
    run.vars[scope + "beta1_power"] = np.float64(np.float32(0.9)); run.vars[scope + "beta2_power"] = np.float64(np.float32(0.999))
    adam = qo.AdamTF1(L["total"])
    p_ref = p.astype(np.float64).copy()
    m_ref = np.zeros(L["total"]); v_ref = np.zeros(L["total"])
    b1, b2, lr, eps = (np.float64(adam.b1), np.float64(adam.b2), np.float64(adam.lr), np.float64(adam.eps))
    b1p, b2p = b1, b2
    for step in range(2):
        new = run.train_step(train_op, feeds)
        # the oracle's gradient at the current parameters + the TF-1 rule in float64
        _, gr, *_ = qo.loss_and_grads(variant, p_ref.astype(np.float32) if step == 0 else p_ref, t, s, s2, a, r, term, None, 0.99, loss_sum) \
            if step == 0 else _grads64(variant, p_ref, t, s, s2, a, y, loss_sum)
        alpha = lr * np.sqrt(1 - b2p) / (1 - b1p)
        m_ref += (gr - m_ref) * (1 - b1); v_ref += (gr * gr - v_ref) * (1 - b2)
        p_ref = p_ref - (m_ref * alpha) / (np.sqrt(v_ref) + eps)
        b1p, b2p = b1p * b1, b2p * b2
        run.vars.update(new)
        for i, name in enumerate(VAR_ORDER):
            o, shp = L[name]
            got = run.vars[prefix + ("Variable" if i == 0 else f"Variable_{i}")]
            ref = p_ref[o:o + int(np.prod(shp))].reshape(shp)
            assert np.abs(got - ref).max() <= 1e-12 + 1e-9 * np.abs(ref).max(), (step, name)
        assert abs(float(run.vars[scope + "beta1_power"]) - b1p) <= 1e-15 and abs(float(run.vars[scope + "beta2_power"]) - b2p) <= 1e-15
    # parameters moved by about lr per step (Adam's normalised first step), i.e. the step really happened
    assert 0.5e-6 < np.abs(p_ref - p.astype(np.float64)).max() < 4e-6


def _grads64(variant, p64, t32, s, s2, a, y, loss_sum):
    """oracle gradient at float64 parameters with the SAME fed y (second optimizer step of the test above)"""
    P = torch.tensor(p64, requires_grad=True)
    q = qo.forward(P, s)
    onehot = torch.nn.functional.one_hot(torch.as_tensor(a.astype(np.int64)), 2).to(torch.float64)
    err = torch.as_tensor(y.astype(np.float64)) - (q * onehot).sum(dim=1)
    loss = (err ** 2).sum() if loss_sum else (err ** 2).mean()
    loss.backward()
    return float(loss.detach()), P.grad.numpy().copy()


def test_interpreter_kernels_against_torch():
    """the NumPy op kernels themselves (Conv2D SAME with stride, its two gradients, MaxPool / MaxPoolGrad) against torch autograd,
    including an odd size where SAME padding is asymmetric (the extra row/column goes at the END in TensorFlow)"""
    rng = np.random.default_rng(0)
    for (H, W, k, s_, ci, co) in [(80, 80, 8, 4, 4, 3), (10, 10, 4, 2, 5, 4), (5, 5, 3, 1, 3, 3), (11, 9, 4, 2, 2, 3)]:
        x = rng.standard_normal((2, H, W, ci)); w = rng.standard_normal((k, k, ci, co))
        y = tg.conv2d(x, w, [1, s_, s_, 1], "SAME")
        OH, pt, pb = tg._same_pad(H, k, s_); OW, pl, pr = tg._same_pad(W, k, s_)
        xt = torch.tensor(np.transpose(x, (0, 3, 1, 2)), requires_grad=True)
        wt = torch.tensor(np.transpose(w, (3, 2, 0, 1)), requires_grad=True)
        yt = torch.nn.functional.conv2d(torch.nn.functional.pad(xt, (pl, pr, pt, pb)), wt, stride=s_)
        assert y.shape == (2, OH, OW, co) and np.abs(y - yt.detach().numpy().transpose(0, 2, 3, 1)).max() <= 1e-10
        dy = rng.standard_normal(y.shape)
        yt.backward(torch.tensor(np.transpose(dy, (0, 3, 1, 2))))
        dw = tg.conv2d_backprop_filter(x, w.shape, dy, [1, s_, s_, 1], "SAME")
        dx = tg.conv2d_backprop_input(x.shape, w, dy, [1, s_, s_, 1], "SAME")
        assert np.abs(dw - wt.grad.numpy().transpose(2, 3, 1, 0)).max() <= 1e-9
        assert np.abs(dx - xt.grad.numpy().transpose(0, 2, 3, 1)).max() <= 1e-9
    x = rng.standard_normal((2, 20, 20, 3)); x[0, 0:2, 0:2, 0] = 0.7                   # a tie: gradient to the FIRST maximum
    y = tg.max_pool(x, [1, 2, 2, 1], [1, 2, 2, 1], "SAME")
    xt = torch.tensor(np.transpose(x, (0, 3, 1, 2)), requires_grad=True)
    yt = torch.nn.functional.max_pool2d(xt, 2)
    assert np.array_equal(y, yt.detach().numpy().transpose(0, 2, 3, 1))
    dy = rng.standard_normal(y.shape)
    dx = tg.max_pool_grad(x, y, dy, [1, 2, 2, 1], [1, 2, 2, 1], "SAME")
    assert dx[0, 0, 0, 0] == dy[0, 0, 0, 0] and dx[0, 0, 1, 0] == 0 and dx[0, 1, 0, 0] == 0 and dx[0, 1, 1, 0] == 0
    assert abs(dx.sum() - dy.sum()) <= 1e-9
    r0, r1 = tg.broadcast_gradient_args([6, 2], [2])
    assert r0.tolist() == [] and r1.tolist() == [0]
