from . import engine, mbd_planner  # noqa: F401
