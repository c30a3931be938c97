from .trees import Strategy, StrategyError, Tree, binary_tree, chain_tree, make_strategy, star_tree  # noqa: F401
from .relay import RelayControl, TreeRole, participants, relay_control, tree_role  # noqa: F401
from .schedule import WorkItem, default_chunk_bytes, slice_bounds, work_items  # noqa: F401
