"""Per-message algorithm plan: the synthesizer's decision of WHICH data-plane variant runs a message of a given size.

The reference executes exactly what it synthesises — one set of trees for every message
(/root/reference/commu.py:272-278 -> /root/reference/csrc/allreduce.cu:723-743). On an NVSwitch box the candidates
are more than trees: LL (flag-in-data, <= 32 KB), one-shot, two-shot, NVLS (in-switch reduction) and the synthesised
strategy trees. ``build_plan`` scores all of them with the alpha-beta model fed by the PROFILED latency / bandwidth
matrices (synth/cost_model.py) over the message-size axis, separately for staged tensors and for tensors inside the
symmetric heap (zero-copy), and compresses the winners into size bands. The bands travel inside the strategy XML
(attributes ``bands`` / ``bands_zc`` of ``<trees>``) so every rank executes the same decision, and the thresholds the
native runtime's own "auto" uses are written to ``topology/tunables.json``. A non-uniform profile (a slow link) moves
the large-message bands from the direct algorithms — bounded by the slowest link every rank must cross — to the trees
that route around it; ``tests/test_synth.py`` checks that flip.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

from ..strategy.trees import Strategy
from .cost_model import LAUNCH_US, LinkModel, direct_times, strategy_time

INF_BYTES = 1 << 62
LL_MAX_BYTES = 32768
# the one-shot kernel has been measured up to 4 MB (where it stops winning on 8 GPUs and ties the two-shot on 2); the model
# would extrapolate it to any size on 2 ranks (same bytes as a two-shot, one staging pass fewer) — not without a measurement
ONE_SHOT_MAX_BYTES = 4 << 20
ALGOS = ("ll", "one_shot", "two_shot", "nvls", "tree")


def ll_time(lm: LinkModel, nbytes: float, ranks: Optional[Sequence[int]] = None) -> float:
    """Flag-in-data one-shot: one kernel, one NVLink store latency, every rank receives (n-1) * 2 * nbytes of lines."""
    rs = list(range(lm.world)) if ranks is None else list(ranks)
    n = max(2, len(rs))
    return 0.6 * LAUNCH_US * 1e-6 + lm.mean_alpha(rs) + 2.0 * (n - 1) * nbytes / (lm.min_bw(rs) * 1e9)


@dataclass
class AlgoPlan:
    """Size bands → algorithm, separately for staged tensors (``bands``) and tensors in the symmetric heap
    (``bands_zc``); serialised into the strategy XML's attributes and consulted per message by
    ``CudaCommu._resolve_algo``."""

    bands: List[Tuple[int, str]] = field(default_factory=list)       # staged: (largest wire bytes of the band, algo)
    bands_zc: List[Tuple[int, str]] = field(default_factory=list)    # tensors in the symmetric heap
    estimates_us: Dict[str, Dict[str, float]] = field(default_factory=dict)   # "<bytes>[zc]" -> algo -> us

    def pick(self, nbytes: int, zero_copy: bool = False, all_active: bool = True, nvls: bool = True,
             ll: bool = True, tree: bool = True) -> str:
        """Algorithm for this message; variants that cannot run (NVLS / LL over a subset, no multicast, no LL
        buffer, no strategy) fall through to the next larger band's choice or to ``auto`` (native policy)."""
        bands = self.bands_zc if zero_copy and self.bands_zc else self.bands
        ok = lambda a: not ((a == "nvls" and not (nvls and all_active)) or (a == "ll" and not (ll and all_active))   # noqa: E731
                            or (a == "tree" and not tree) or (a == "one_shot" and zero_copy))
        for i, (mx, algo) in enumerate(bands):
            if nbytes <= mx:
                for _, a in bands[i:]:
                    if ok(a) and not (a == "ll" and nbytes > LL_MAX_BYTES):
                        return a
                break
        return "auto"

    @staticmethod
    def _enc(bands) -> str:
        return ",".join(f"{a}:{'inf' if mx >= INF_BYTES else mx}" for mx, a in bands)

    @staticmethod
    def _dec(text: str) -> List[Tuple[int, str]]:
        out = []
        for part in (text or "").split(","):
            if ":" not in part:
                continue
            a, mx = part.split(":", 1)
            if a in ALGOS:
                out.append((INF_BYTES if mx == "inf" else int(mx), a))
        return out

    def to_attrs(self) -> Dict[str, str]:
        return {"bands": self._enc(self.bands), "bands_zc": self._enc(self.bands_zc)}

    @classmethod
    def from_attrs(cls, attrs: Dict[str, str]) -> Optional["AlgoPlan"]:
        b, z = cls._dec(attrs.get("bands", "")), cls._dec(attrs.get("bands_zc", ""))
        return cls(b, z) if b or z else None

    def tunables(self) -> Dict[str, int]:
        """Thresholds for the native runtime's own policy (CommContext::pick_algo), derived from the same bands."""
        t: Dict[str, int] = {}
        prev = 0
        for mx, a in self.bands:
            if a == "one_shot":
                t["one_shot_max_bytes"] = min(mx, 64 << 20)
            if a == "nvls" and "nvls_min_bytes" not in t:
                t["nvls_min_bytes"] = prev + 1 if prev else 0
            prev = mx
        if "one_shot_max_bytes" not in t:
            t["one_shot_max_bytes"] = 0
        t["ll_max_bytes"] = max([mx for mx, a in self.bands if a == "ll"] + [0])
        if not any(a == "nvls" for _, a in self.bands + self.bands_zc):
            t["nvls_min_bytes"] = INF_BYTES            # never
        return t


# the tree kernel pipelines at this granularity whatever chunk size the API asks for (tunable tree_chunk_max_bytes);
# the cost model is calibrated with it (tests/test_synth.py::test_tree_cost_model_against_the_measured_tree_kernel)
TREE_DEVICE_CHUNK = 256 << 10


def build_plan(lm: LinkModel, strategy: Optional[Strategy] = None, chunk_bytes: float = TREE_DEVICE_CHUNK, nvls: bool = True,
               nvls_bw_gbs: Optional[float] = None, ll: bool = True,
               sizes: Optional[Sequence[int]] = None) -> AlgoPlan:
    sizes = list(sizes) if sizes is not None else [1 << p for p in range(8, 32)]
    plan = AlgoPlan()
    for zc, dst in ((False, plan.bands), (True, plan.bands_zc)):
        winners = []
        for nb in sizes:
            t = direct_times(lm, float(nb), None, nvls, nvls_bw_gbs or None, zero_copy=zc)
            if zc or nb > ONE_SHOT_MAX_BYTES:
                t.pop("one_shot", None)
            if ll and nb <= LL_MAX_BYTES:
                t["ll"] = ll_time(lm, float(nb))
            if strategy is not None and strategy.trees:
                t["tree"] = strategy_time(strategy, lm, float(nb), min(chunk_bytes, TREE_DEVICE_CHUNK))
            best = min(t, key=t.get)
            winners.append((nb, best))
            plan.estimates_us[f"{nb}{'zc' if zc else ''}"] = {k: v * 1e6 for k, v in t.items()}
        for i, (nb, a) in enumerate(winners):
            last = i == len(winners) - 1
            if last or winners[i + 1][1] != a:
                dst.append((INF_BYTES if last else nb, a))
    return plan
