from .synthesizer import Synthesizer  # noqa: F401
from .partrees import ParTrees, DEFAULT_CHUNK  # noqa: F401
from .solver import Solver, SolverError  # noqa: F401
from .cost_model import LinkModel, direct_times, pick_algorithm, strategy_time, best_chunk_bytes, crossover_bytes  # noqa: F401
