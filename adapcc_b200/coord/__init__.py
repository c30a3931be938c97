from .server import Coordinator, make_server  # noqa: F401
from .client import Controller, Hooker  # noqa: F401
