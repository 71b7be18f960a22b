from . import engine, mbd_planner, path_integral  # noqa: F401
