"""tdmpc_b200: B200-native (sm_100a) planner for TD-MPC -- drop-in for `TDMPC.plan()`."""
from .told import DSSM, DSSMRnn, TOLD, linear_schedule  # noqa: F401
from .planner import Planner, TDMPC, TDMPCSIM, TdICemSimDssm, TdICemSimMlp, accelerate, agent_kind, icem_sizes  # noqa: F401
from .collector import BatchedEpisodes, PlanLoop  # noqa: F401,E402
from .replay import ReplayBuffer  # noqa: F401,E402
