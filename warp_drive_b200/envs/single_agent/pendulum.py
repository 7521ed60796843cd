"""Import path of the reference's example_envs/single_agent/classic_control/pendulum/pendulum.py;
the classes live in classic_control.py."""
from warp_drive_b200.envs.single_agent.classic_control import (  # noqa: F401
    ClassicControlPendulumEnv,
    CUDAClassicControlPendulumEnv,
)
