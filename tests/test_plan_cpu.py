"""Plan-compiler logic on CPU: the compiled node list, interpreted with torch ops (tests/plan_interp.py),
must reproduce the oracle.  Covers fusions done at plan time (concat-by-offset, upsample folding, realign at
low resolution, stacked q/k/v weights, qkv channel re-ordering) for every graph family of BASELINE.json."""
import pytest
import torch

from cases import FORWARD_CASES, MODULE_CASES
from helpers import golden, load_cfg, rel_l2, rel_max
from plan_interp import Interp
from test_oracle_golden import module_case_io

from fce_yolo_b200.plan import compile_model, compile_module
from fce_yolo_b200.tasks import DetectionModel
from fce_yolo_b200.weights import load_synthetic, synth_images
from oracle import fce_oracle as O


@pytest.mark.parametrize("name", ["n_fce_64", "s_coordatt_64", "s_cca_bicca8_64", "m_bifpn_64", "x_fce_64",
                                  "n_stock_64"])
def test_model_plan_matches_oracle_and_golden(name):
    case = FORWARD_CASES[name]
    cfg, scale = load_cfg(case)
    model = DetectionModel(cfg, scale=scale).fuse().eval()
    sd = load_synthetic(model, case["seed"])
    x = synth_images(case["img_seed"], case["batch"], case["size"], case["size"])
    plan = compile_model(model, case["batch"], case["size"], case["size"], "fp32", torch.device("cpu"))
    it = Interp(plan, reuse_memory=True)
    it.input_tensor().copy_(x)
    it.run()
    y, raw = it.outputs()
    (yo, rawo), ys = O.forward(cfg, scale, sd, x, keep_layers=True)
    assert rel_max(y, yo) < 1e-4
    for a, b in zip(raw, rawo):
        assert rel_max(a, b) < 1e-4
    g = golden("fwd_" + name)
    assert rel_max(y, g["y"]) < 1e-4


def test_memory_reuse_shrinks_arena():
    case = FORWARD_CASES["n_fce_64"]
    cfg, scale = load_cfg(case)
    model = DetectionModel(cfg, scale=scale).fuse().eval()
    from fce_yolo_b200.engine import assign_offsets

    p1 = compile_model(model, 2, 64, 64, "bf16", torch.device("cpu"))
    p2 = compile_model(model, 2, 64, 64, "bf16", torch.device("cpu"))
    assert assign_offsets(p1, reuse=True) < 0.6 * assign_offsets(p2, reuse=False)


def test_unfused_model_folds_bn_at_plan_time():
    """A model that still carries BatchNorm (e.g. a freshly loaded checkpoint) compiles to the same plan as
    its fused twin: BN is folded when weights are packed (reference torch_utils.py:237-267)."""
    cfg, scale = load_cfg(FORWARD_CASES["n_fce_64"])
    m = DetectionModel(cfg, scale=scale).eval()
    g = torch.Generator().manual_seed(3)
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.copy_(torch.randn(mod.num_features, generator=g) * 0.1)
            mod.running_var.copy_(torch.rand(mod.num_features, generator=g) + 0.5)
            mod.weight.data.copy_(torch.rand(mod.num_features, generator=g) + 0.5)
            mod.bias.data.copy_(torch.randn(mod.num_features, generator=g) * 0.1)
    x = synth_images(5, 1, 64, 64)
    outs = []
    for fused in (False, True):
        if fused:
            m.fuse()
        it = Interp(compile_model(m, 1, 64, 64, "fp32", torch.device("cpu")))
        it.input_tensor().copy_(x)
        it.run()
        outs.append(it.outputs()[0].clone())
    assert rel_max(outs[0], outs[1]) < 1e-5


@pytest.mark.parametrize("name", list(MODULE_CASES))
def test_module_plan_matches_golden(name):
    case = MODULE_CASES[name]
    mod, sd, xs = module_case_io(case)
    plan = compile_module(mod.eval(), [tuple(t.shape) for t in xs], "fp32", torch.device("cpu"))
    it = Interp(plan, reuse_memory=False)
    for i, t in enumerate(xs):
        it.nchw(plan.inputs[i]).copy_(t)
    it.run()
    y = it.outputs()
    assert rel_max(y, golden("mod_" + name)["y"]) < 1e-5


def test_cca_bad_heads_fails_at_build():
    from fce_yolo_b200 import modules as M

    with pytest.raises(RuntimeError):
        M.CoordCrossAtt(512, 512, 22, 2)  # mip = 23, heads = 2: the reference dies in forward (fce_block.py:166)


def test_fused_decode_plan_structure_and_values():
    """compile_model(fuse_decode=True) in bf16: Detect's last convs become fce_conv2d_detect nodes placed LAST (they
    are the short tail that waits for the previous call's NMS in overlap mode), there is no decode node and no logit
    map; the interpreted plan gives the same predictions as the unfused plan."""
    case = FORWARD_CASES["n_fce_64"]
    cfg, scale = load_cfg(case)
    model = DetectionModel(cfg, scale=scale).fuse().eval()
    load_synthetic(model, case["seed"])
    x = synth_images(case["img_seed"], case["batch"], case["size"], case["size"])
    plans = {f: compile_model(model, case["batch"], case["size"], case["size"], "bf16", torch.device("cpu"), fuse_decode=f)
             for f in (False, True)}
    fns = [n.fn for n in plans[True].nodes]
    assert "fce_detect_decode" not in fns and fns[-6:] == ["fce_conv2d_detect"] * 6
    assert plans[True].outputs["raw"] == [] and len(plans[False].outputs["raw"]) == 3
    assert [n.fn for n in plans[False].nodes].count("fce_detect_decode") == 1
    # fp32 plans keep the unfused route (the fused epilogue exists on the tensor-core path only)
    assert "fce_conv2d_detect" not in [n.fn for n in compile_model(model, 1, 64, 64, "fp32", torch.device("cpu"),
                                                                    fuse_decode=True).nodes]
    ys = {}
    for f, plan in plans.items():
        it = Interp(plan, reuse_memory=True)
        it.input_tensor().copy_(x)
        it.run()
        ys[f] = it.outputs()[0].float()
    assert rel_max(ys[True], ys[False]) < 1e-3  # same bf16 activations; only the logits skip their fp32 round trip


@pytest.mark.parametrize("name", ["n_fce_64", "s_coordatt_64", "m_bifpn_64"])
def test_bifpn_fused_into_realign_convs(name):
    """bf16 plans fold BiFPN_Concat's weighted sum (fce_block.py:55-63) into the epilogue of its full-resolution
    realign convs (fce_conv_desc.out_scale / res_scale / res_up).  Same predictions as the separate
    fce_bifpn_fuse launches up to bf16 rounding of the intermediates; nodes without a fusable conv keep the kernel."""
    from fce_yolo_b200.plan import Plan

    case = FORWARD_CASES[name]
    cfg, scale = load_cfg(case)
    model = DetectionModel(cfg, scale=scale).fuse().eval()
    load_synthetic(model, case["seed"])
    x = synth_images(case["img_seed"], case["batch"], case["size"], case["size"])
    ys, counts = {}, {}
    for fused in (False, True):
        Plan.FUSED_BIFPN, Plan.FUSED_SUM = fused, False  # FUSED_SUM: layer outputs of the BiFPN nodes are read back below
        try:
            plan = compile_model(model, case["batch"], case["size"], case["size"], "bf16", torch.device("cpu"))
        finally:
            Plan.FUSED_BIFPN = Plan.FUSED_SUM = True
        counts[fused] = sum(1 for n in plan.nodes if n.fn == "fce_bifpn_fuse")
        it = Interp(plan, reuse_memory=False)  # layer outputs are read back after the run
        it.input_tensor().copy_(x)
        it.run()
        bif = [i for i, m in enumerate(model.model) if type(m).__name__ == "BiFPN_Concat"]
        ys[fused] = (it.outputs()[0].float(), [it.nchw(plan.layer_out[i]).float() for i in bif])
    assert counts[False] == 4 and counts[True] <= counts[False]
    if name != "m_bifpn_64":  # m scale: every input but one is an Identity realign
        assert counts[True] < counts[False]
    # the first BiFPN node sees identical inputs in both plans: only the rounding of the intermediates differs
    assert rel_max(ys[True][1][0], ys[False][1][0]) < 1e-2
    for a, b in zip(ys[True][1], ys[False][1]):
        assert rel_l2(a, b) < 2e-2
    assert rel_l2(ys[True][0], ys[False][0]) < 5e-2  # end to end: two bf16 evaluation orders of the same graph


@pytest.mark.parametrize("name", ["m_bifpn_64", "n_fce_64"])
def test_bifpn_sum_folded_into_consumer(name):
    """Two-input BiFPN nodes with identity realigns (m scale: SURVEY A.1-3) never store their weighted sum: the C3k2.cv1
    behind them runs as (w0 W) a [at a's resolution] + (w1 W) b with the first product joining the second conv's
    accumulator (fce_conv_desc.weighted == 2).  Same predictions as the fce_bifpn_fuse route up to bf16 rounding."""
    from fce_yolo_b200.plan import Plan

    case = FORWARD_CASES[name]
    cfg, scale = load_cfg(case)
    model = DetectionModel(cfg, scale=scale).fuse().eval()
    load_synthetic(model, case["seed"])
    x = synth_images(case["img_seed"], case["batch"], case["size"], case["size"])
    ys, fuse_nodes, folded = {}, {}, {}
    for fused in (False, True):
        Plan.FUSED_SUM = fused
        try:
            plan = compile_model(model, case["batch"], case["size"], case["size"], "bf16", torch.device("cpu"))
        finally:
            Plan.FUSED_SUM = True
        fuse_nodes[fused] = sum(1 for n in plan.nodes if n.fn == "fce_bifpn_fuse")
        folded[fused] = sum(1 for n in plan.nodes if n.fn == "fce_conv2d" and n.desc.weighted == 2)
        it = Interp(plan, reuse_memory=True)
        it.input_tensor().copy_(x)
        it.run()
        ys[fused] = it.outputs()[0].float()
    assert folded[False] == 0
    if name == "m_bifpn_64":
        assert folded[True] == 2 and fuse_nodes[True] == fuse_nodes[False] - 2  # three-input / same-resolution nodes keep the kernel
    else:
        assert folded[True] == fuse_nodes[False] - fuse_nodes[True]
    assert rel_l2(ys[True], ys[False]) < 5e-2
    # fp32 plans keep the separate launch
    assert all(n.desc.weighted != 2 for n in compile_model(model, 1, 64, 64, "fp32", torch.device("cpu")).nodes
               if n.fn == "fce_conv2d")


def test_detect_branches_and_dependencies():
    """Branching plans (small workloads): the 2 * nl Detect conv chains, C3k's cv2 and BiFPN's realign convs go to side
    branches (Node.stream > 0).  Plan.dependencies() must order every Detect chain after the neck node that produces
    its level (and after nothing later), chain nodes after their predecessor, the six tails only after their own chain
    (they write disjoint parts of y), and the NMS after all tails; C3k.cv2 must not be ordered against the bottleneck
    chain that fills the other half of the concat buffer; buffers touched by a branch are excluded from the
    sequential lifetime packing."""
    from fce_yolo_b200.plan import Plan

    case = FORWARD_CASES["n_fce_64"]
    cfg, scale = load_cfg(case)
    model = DetectionModel(cfg, scale=scale).fuse().eval()
    saved, Plan.HEAD_STREAMS = Plan.HEAD_STREAMS, True
    saved_c3k, Plan.FUSED_C3K_IN = Plan.FUSED_C3K_IN, False  # the unfused C3k: cv2 is a branch of its own
    try:
        plan = compile_model(model, 2, 64, 64, "bf16", torch.device("cpu"), fuse_decode=True,
                             nms=dict(conf=0.25, iou=0.7, max_det=300))
        Plan.FUSED_C3K_IN = True
        fused = compile_model(model, 2, 64, 64, "bf16", torch.device("cpu"), fuse_decode=True,
                              nms=dict(conf=0.25, iou=0.7, max_det=300))
    finally:
        Plan.HEAD_STREAMS = saved
        Plan.FUSED_C3K_IN = saved_c3k
    # fused C3k input convs (the default): one stacked launch per C3k feeds both the bottleneck chain and cv3
    fdeps = fused.dependencies()
    stacked = [i for i, n in enumerate(fused.nodes) if n.tag.endswith(".cv2+cv1")]
    assert stacked and len(fused.nodes) == len(plan.nodes) - len(stacked)
    for i in stacked:
        base = fused.nodes[i].tag[:-len(".cv2+cv1")]
        chain = [j for j, n in enumerate(fused.nodes) if n.tag.startswith(base + ".m.")]
        cv3 = next(j for j, n in enumerate(fused.nodes) if n.tag in (base + ".cv3", base + ".cv3+cv2"))  # alone, or chained into C3k2.cv2
        assert i in fdeps[chain[0]] and i in fdeps[cv3] and chain[-1] in fdeps[cv3]
    nodes, deps = plan.nodes, plan.dependencies()
    by_stream = {}
    for i, n in enumerate(nodes):
        by_stream.setdefault(n.stream, []).append(i)
    head = {k: v for k, v in by_stream.items() if k and nodes[v[0]].tag.startswith("model.25")}
    assert len(head) == 6
    level_of = {"0": "model.18", "1": "model.21", "2": "model.24"}
    for k, chain in head.items():
        tag = nodes[chain[0]].tag  # model.25.cv2.<level>.0 or model.25.cv3.<level>.0.0
        n_dwpw = sum(nodes[i].fn == "fce_dwpw_conv" for i in chain)  # a DWConv + 1x1 block the fused kernel takes is ONE node
        assert nodes[chain[-1]].fn == "fce_conv2d_detect" and len(chain) == (3 if ".cv2." in tag else 5 - n_dwpw)
        assert n_dwpw == (0 if ".cv2." in tag else 1)  # n scale: 64 -> 80 fuses, the 80-channel depthwise does not
        (src,) = deps[chain[0]]
        assert nodes[src].stream == 0 and nodes[src].tag.startswith(level_of[tag.split(".")[3]])
        for a, b in zip(chain, chain[1:]):
            assert deps[b] == {a}
        for i in chain:
            for v in nodes[i].reads + nodes[i].writes:
                assert v.buf.persistent
    tails = {chain[-1] for chain in head.values()}
    assert nodes[-1].fn == "fce_nms" and tails <= deps[len(nodes) - 1]
    # C3k: cv2 (branch) and the bottleneck chain write different halves of one buffer -> no edge between them; cv3 waits
    # for both
    c3k_cv2 = [i for i, n in enumerate(nodes) if n.stream and n.tag.endswith(".m.0.cv2") and "model.25" not in n.tag]
    assert c3k_cv2
    for i in c3k_cv2:
        base = nodes[i].tag[:-len(".cv2")]
        chain = [j for j, n in enumerate(nodes) if n.tag.startswith(base + ".m.")]
        cv3 = next(j for j, n in enumerate(nodes) if n.tag in (base + ".cv3", base + ".cv3+cv2"))
        assert chain and not (set(chain) & deps[i]) and i in deps[cv3] and chain[-1] in deps[cv3]
    # main-stream nodes wait for a branch only when they consume its output
    for i in by_stream[0][:-1]:
        for j in deps[i]:
            if nodes[j].stream:
                assert any(Plan._overlap(r, w) for r in nodes[i].reads for w in nodes[j].writes) or \
                    any(Plan._overlap(a, b) for a in nodes[i].writes for b in nodes[j].reads + nodes[j].writes)
    # a sub-range (the overlap mode's second graph: the six tails) has no internal edges
    k = min(i for i, n in enumerate(nodes) if n.fn == "fce_conv2d_detect")
    sub = plan.dependencies(k, len(nodes) - 1)
    assert all(not d for d in sub.values())
    # large workloads and unfused plans stay on one stream
    assert all(n.stream == 0 for n in compile_model(model, 2, 64, 64, "bf16", torch.device("cpu")).nodes)
    big = compile_model(model, 64, 640, 640, "bf16", torch.device("cpu"), fuse_decode=True)
    assert all(n.stream == 0 for n in big.nodes)


@pytest.mark.parametrize("name", ["n_fce_64", "s_coordatt_64"])
def test_branch_schedules_are_equivalent(name):
    """The executor orders side branches with events derived from Plan.dependencies(); everything else may run in
    any interleaving.  Interpret the branching plan in RANDOM legal orders (per-branch order kept, cross-branch edges
    respected) on the SAME packed arena the executor uses (lifetime-recycled offsets): a missing edge or a buffer
    recycled under a running branch would change the predictions."""
    import random

    from fce_yolo_b200.plan import Plan

    case = FORWARD_CASES[name]
    cfg, scale = load_cfg(case)
    model = DetectionModel(cfg, scale=scale).fuse().eval()
    load_synthetic(model, case["seed"])
    x = synth_images(case["img_seed"], case["batch"], case["size"], case["size"])
    saved, Plan.HEAD_STREAMS = Plan.HEAD_STREAMS, True
    try:
        plan = compile_model(model, case["batch"], case["size"], case["size"], "bf16", torch.device("cpu"),
                             fuse_decode=True)
    finally:
        Plan.HEAD_STREAMS = saved
    nodes, deps = plan.nodes, plan.dependencies()
    assert max(n.stream for n in nodes) >= 6

    def run(order):
        it = Interp(plan, reuse_memory=True)
        it.input_tensor().copy_(x)
        for i in order:
            n = nodes[i]
            getattr(it, "_" + n.fn)(n.desc, n.ptrs)
        return it.outputs()[0].float().clone()

    ref = run(range(len(nodes)))
    rng = random.Random(7)
    for trial in range(4):
        queues = {}
        for i, n in enumerate(nodes):
            queues.setdefault(n.stream, []).append(i)
        done, order = set(), []
        while len(order) < len(nodes):
            ready = [k for k, q in queues.items() if q and all(j in done for j in deps[q[0]] if nodes[j].stream != k)]
            assert ready, "dependency cycle"
            # bias towards the side branches so that they run as EARLY as the edges allow
            k = rng.choice([s for s in ready if s] or ready) if trial % 2 == 0 else rng.choice(ready)
            i = queues[k].pop(0)
            order.append(i)
            done.add(i)
        assert order != list(range(len(nodes)))
        assert torch.equal(run(order), ref), trial


@pytest.mark.parametrize("name", ["n_fce_64", "s_coordatt_64", "s_cca_bicca8_64", "m_bifpn_64", "x_fce_64", "n_stock_64"])
def test_predict_plan_bf16_all_families(name):
    """The predict plan exactly as the Predictor compiles it (bf16, fused Detect epilogues, BiFPN folded into realign
    convs, CoordAtt MLP, graph branches) for every graph family of BASELINE.json, interpreted on CPU: predictions must
    agree with the fp32 oracle like any bf16 evaluation does and with the unfused bf16 plan more closely."""
    from fce_yolo_b200.plan import Plan

    case = FORWARD_CASES[name]
    cfg, scale = load_cfg(case)
    model = DetectionModel(cfg, scale=scale).fuse().eval()
    sd = load_synthetic(model, case["seed"])
    x = synth_images(case["img_seed"], case["batch"], case["size"], case["size"])
    yo, _ = O.forward(cfg, scale, sd, x)
    ys = {}
    for fused in (False, True):
        saved = (Plan.HEAD_STREAMS, Plan.FUSED_BIFPN, Plan.FUSED_COORDATT_MLP)
        Plan.HEAD_STREAMS, Plan.FUSED_BIFPN, Plan.FUSED_COORDATT_MLP = fused, fused, fused
        try:
            plan = compile_model(model, case["batch"], case["size"], case["size"], "bf16", torch.device("cpu"),
                                 fuse_decode=fused, nms=dict(conf=0.25, iou=0.7, max_det=300))
        finally:
            Plan.HEAD_STREAMS, Plan.FUSED_BIFPN, Plan.FUSED_COORDATT_MLP = saved
        fns = [n.fn for n in plan.nodes]
        assert fns[-1] == "fce_nms"
        assert ("fce_conv2d_detect" in fns) == fused and ("fce_detect_decode" in fns) == (not fused)
        it = Interp(plan, reuse_memory=True)
        it.input_tensor().copy_(x)
        for n in plan.nodes[:-1]:  # the NMS node is covered by the GPU / oracle NMS tests
            getattr(it, "_" + n.fn)(n.desc, n.ptrs)
        ys[fused] = it.outputs()[0].float().clone()
    # vs the fp32 oracle: the end-to-end drift of ANY bf16 evaluation of these random-weight graphs (the reference's own
    # bf16 forward reaches 0.16 relL2; saturated class logits flip individual scores) - a loose sanity bound
    for y in ys.values():
        assert rel_l2(y[:, :4], yo[:, :4]) < 5e-2
        assert rel_l2(y[:, 4:], yo[:, 4:]) < 0.25
    # fused vs unfused plan: two bf16 evaluation orders of the same graph
    assert rel_l2(ys[True][:, :4], ys[False][:, :4]) < 3e-3   # measured 3e-8 .. 4e-4
    assert rel_l2(ys[True][:, 4:], ys[False][:, 4:]) < 3e-2   # measured 0 .. 5e-3


@pytest.mark.parametrize("name", ["m_bifpn_64", "s_coordatt_64", "n_fce_64"])
def test_fused_producer_nodes_in_u8_plans(name):
    """uint8-input bf16 plans: the first two convs become ONE fce_stem2_conv node where the kernel takes the stem width (32 /
    64 channels: s, m, l scale - not the 16-channel n-scale stem) and every DWConv + 1x1 block of Detect's class branch ONE
    fce_dwpw_conv node where it takes the channel counts; the interpreted plan gives the predictions of the plan with the
    two-launch routes (same layer arithmetic, interpreted in fp32 with bf16 storage)."""
    from fce_yolo_b200.plan import Plan

    case = FORWARD_CASES[name]
    cfg, scale = load_cfg(case)
    model = DetectionModel(cfg, scale=scale).fuse().eval()
    load_synthetic(model, case["seed"])
    B, S = case["batch"], case["size"]
    x = (synth_images(case["img_seed"], B, S, S) * 255).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous()
    saved = (Plan.FUSED_STEM2, Plan.FUSED_DWPW)
    plans = {}
    try:
        for fused in (False, True):
            Plan.FUSED_STEM2 = Plan.FUSED_DWPW = fused
            plans[fused] = compile_model(model, B, S, S, "bf16", torch.device("cpu"), input_u8=True)
    finally:
        Plan.FUSED_STEM2, Plan.FUSED_DWPW = saved
    fns = {f: [n.fn for n in p.nodes] for f, p in plans.items()}
    assert "fce_stem2_conv" not in fns[False] and "fce_dwpw_conv" not in fns[False]
    stem_width = model.model[0].conv.out_channels
    assert ("fce_stem2_conv" in fns[True]) == (stem_width in (32, 64))
    n_dwpw = fns[True].count("fce_dwpw_conv")
    assert 1 <= n_dwpw <= 6
    assert len(fns[True]) == len(fns[False]) - n_dwpw - ("fce_stem2_conv" in fns[True])
    if "fce_stem2_conv" in fns[True]:
        assert fns[True][0] == "fce_stem2_conv" and "fce_stem_conv" not in fns[True]
    ys = {}
    for f, plan in plans.items():
        it = Interp(plan, reuse_memory=True)
        it.input_tensor().copy_(x)
        it.run()
        ys[f] = it.outputs()[0].float()
    assert rel_max(ys[True], ys[False]) < 1e-5  # the interpreter runs the same fp32 arithmetic with the same bf16 roundings


@pytest.mark.parametrize("name", ["m_bifpn_64", "n_fce_64", "x_fce_64", "s_coordatt_64"])
def test_c3k_tail_chained_into_c3k2_cv2(name):
    """bf16 plans: where a C3k2 block's last inner block is a C3k with 64 / 128 hidden channels and cv2 has a multiple of 128
    outputs, C3k.cv3 and C3k2.cv2 become ONE fce_conv1x1_chain node and the concat buffer loses the slot of the C3k output
    (m scale: layers 2, 4, 16; n scale: 6, 8, 22; s scale: 6; wider C3k blocks, the x-scale widths 96 / 192 and the
    plain-bottleneck blocks stay as they are); the interpreted plan gives the predictions of the two-launch plan."""
    from fce_yolo_b200.plan import Plan

    case = FORWARD_CASES[name]
    cfg, scale = load_cfg(case)
    model = DetectionModel(cfg, scale=scale).fuse().eval()
    load_synthetic(model, case["seed"])
    B, S = case["batch"], case["size"]
    x = synth_images(case["img_seed"], B, S, S)
    saved = Plan.FUSED_C3K_TAIL
    plans = {}
    try:
        for fused in (False, True):
            Plan.FUSED_C3K_TAIL = fused
            plans[fused] = compile_model(model, B, S, S, "bf16", torch.device("cpu"))
    finally:
        Plan.FUSED_C3K_TAIL = saved
    tags = {f: [n.tag for n in p.nodes] for f, p in plans.items()}
    chains = [n for n in plans[True].nodes if n.fn == "fce_conv1x1_chain"]
    assert not any(n.fn == "fce_conv1x1_chain" for n in plans[False].nodes)
    expect = 0
    for i, m in enumerate(model.model):
        if type(m).__name__ == "C3k2" and len(m.m) and type(m.m[-1]).__name__ == "C3k":
            expect += m.c in (64, 128) and m.cv2.conv.out_channels % 128 == 0
    assert len(chains) == expect and len(tags[True]) == len(tags[False]) - expect
    assert expect == {"m_bifpn_64": 3, "n_fce_64": 3, "s_coordatt_64": 1, "x_fce_64": 0}[name]
    for n in chains:
        d, base = n.desc, n.tag[:-len(".cv3+cv2")]
        assert base + ".cv3" in tags[False] and base + ".cv3" not in tags[True]
        assert (d.c1, d.cm) == (d.cm, d.c1) and d.c2 % d.cm == 0 and d.x2_pitch == d.c2  # no slot for t in the concat buffer
        x1, x2 = n.ptrs[0], n.ptrs[3]
        assert (x1.C, x1.pitch, x2.C, x2.pitch) == (d.c1, d.x1_pitch, d.c2, d.x2_pitch) and n.reads == [x1, x2]
    ys = {}
    for f, plan in plans.items():
        it = Interp(plan, reuse_memory=True)
        it.input_tensor().copy_(x.permute(0, 2, 3, 1) if it.input_tensor().shape[-1] == 3 else x)
        it.run()
        ys[f] = it.outputs()[0].float()
    assert rel_max(ys[True], ys[False]) < 1e-5  # the interpreter runs the same fp32 arithmetic with the same bf16 roundings


def test_capture_keeps_the_cycle_collector_out(monkeypatch):
    """Executor._capture: a full collection BEFORE the capture window, the collector off inside it (finalizing a dead executor
    there would destroy its CUDA graphs - forbidden during capture, it invalidates the open one), back on afterwards, also
    when the capture raises.  CUDA-free: torch.cuda.CUDAGraph / torch.cuda.graph are replaced by recorders."""
    import gc

    from fce_yolo_b200 import engine

    log = []

    class FakeGraph:
        pass

    class FakeCtx:
        def __init__(self, g, stream=None):
            log.append(("ctx", isinstance(g, FakeGraph), stream))

        def __enter__(self):
            log.append(("enter", gc.isenabled()))

        def __exit__(self, *exc):
            log.append(("exit", gc.isenabled()))

    monkeypatch.setattr(torch.cuda, "CUDAGraph", FakeGraph)
    monkeypatch.setattr(torch.cuda, "graph", FakeCtx)
    collected = []
    monkeypatch.setattr(engine.gc, "collect", lambda *a: collected.append(gc.isenabled()) or 0)
    ex = object.__new__(engine.Executor)
    ex._capture_stream = lambda: "cap"
    ex._launch_branched = lambda lo, hi: log.append(("launch", lo, hi, gc.isenabled()))
    assert gc.isenabled()
    g = ex._capture(3, 7)
    assert isinstance(g, FakeGraph) and collected == [True] and gc.isenabled()
    assert log == [("ctx", True, "cap"), ("enter", False), ("launch", 3, 7, False), ("exit", False)]

    def boom(lo, hi):
        raise RuntimeError("launch failed")

    ex._launch_branched = boom
    with pytest.raises(RuntimeError):
        ex._capture(0, 1)
    assert gc.isenabled()
    gc.disable()  # a host application that runs with the collector off keeps it off
    try:
        ex._launch_branched = lambda lo, hi: None
        ex._capture(0, 1)
        assert not gc.isenabled()
    finally:
        gc.enable()
