"""Driver contract checks that do not need a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_reports_unavailable():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1",
                        "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    rec = json.loads(r.stdout.strip().splitlines()[-1])
    assert rec["impl"] == "reference" and "unavailable" in rec and len(rec["unavailable"]) > 20


def test_graft_entry_build_and_exports():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g

    g.build()
    from adapcc_b200.runtime.native import lib_path, load_library

    lib = load_library(build_if_missing=False)
    assert os.path.exists(lib_path())
    for sym in ("initThreads", "exitThreads", "allreduce", "reduce", "boardcast", "updateActive", "adapcc_allreduce",
                "adapcc_tree_collective", "adapcc_tree_relay_persistent", "adapcc_profile_links",
                "adapcc_detect_topology", "adapcc_fused_adamw", "adapcc_fused_ce", "adapcc_moe_exchange",
                "adapcc_alltoall", "adapcc_ln_fwd", "adapcc_pool_alloc"):
        assert hasattr(lib, sym), sym


def test_shipped_strategies_parse_and_validate():
    from adapcc_b200.strategy import Strategy

    d = os.path.join(ROOT, "strategy")
    files = [f for f in os.listdir(d) if f.endswith(".xml")]
    assert len(files) >= 10
    for f in files:
        s = Strategy.from_file(os.path.join(d, f), None)
        assert s.trees
        world = len(s.ranks())
        s.validate(world)
