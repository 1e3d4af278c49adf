"""Static checks of bench.py's multi-rank control flow (no GPU): code that only rank 0 executes must not contain a
collective or a call that issues one - the other ranks would never join it (this hung a 4-GPU run once)."""
import ast
import os

from helpers import ROOT

COLLECTIVE_CALLS = {"step_device", "gather_detections", "gather", "barrier", "all_reduce", "all_gather", "all_gather_into_tensor",
                    "broadcast", "gather_stats_to_rank0", "init_process_group", "destroy_process_group"}


def _calls(node):
    for n in ast.walk(node):
        if isinstance(n, ast.Call):
            f = n.func
            yield f.attr if isinstance(f, ast.Attribute) else getattr(f, "id", "")


def _mentions_rank(test):
    return any(isinstance(n, ast.Name) and n.id == "rank" for n in ast.walk(test))


def test_rank_conditional_code_has_no_collectives():
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    checked = 0
    for node in ast.walk(tree):
        if isinstance(node, ast.If) and _mentions_rank(node.test):
            # `if rank != 0: return` in the CPU reference arm is fine: nothing after it is collective in that arm
            for stmt in node.body + node.orelse:
                bad = COLLECTIVE_CALLS & set(_calls(stmt))
                assert not bad, f"bench.py line {stmt.lineno}: {sorted(bad)} inside rank-conditional code"
            checked += 1
    assert checked >= 4


def test_every_rank_reaches_the_same_collectives():
    """The barriers / all_reduce of the GPU arm sit at function level of measure() (which main() calls at function level
    for the judged workload), not under any condition on rank."""
    src = open(os.path.join(ROOT, "bench.py")).read()
    tree = ast.parse(src)
    meas = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "measure")
    top = [c for stmt in meas.body if not isinstance(stmt, (ast.If, ast.FunctionDef)) for c in _calls(stmt)]
    assert top.count("barrier") >= 5
    main = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "main")
    top_main = [c for stmt in main.body if not isinstance(stmt, (ast.If, ast.FunctionDef)) for c in _calls(stmt)]
    assert top_main.count("measure") == 1
