"""tdmpc_b200: B200-native (sm_100a) planner for TD-MPC -- drop-in for `TDMPC.plan()`."""
from .told import DSSM, TOLD, linear_schedule  # noqa: F401
from .planner import Planner, TDMPC, TdICemSimMlp, accelerate, icem_sizes  # noqa: F401
