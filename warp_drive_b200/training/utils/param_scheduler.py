"""Constant / piecewise-linear parameter schedules (learning rate, loss coefficients).
Same config syntax as warp_drive/training/utils/param_scheduler.py:16-75: a number, or a
list of [timestep, value] knots with increasing timesteps, linearly interpolated."""


class ParamScheduler:
    def __init__(self, schedule):
        if isinstance(schedule, (int, float)):
            self.type = "constant"
        elif isinstance(schedule, list):
            self.type = "piecewise_linear"
            assert all(isinstance(k, (list, tuple)) and len(k) == 2 for k in schedule), (
                "each schedule entry must be [timestep, value]")
            times = [k[0] for k in schedule]
            assert times == sorted(times), "schedule timesteps must be increasing"
        else:
            raise NotImplementedError(f"unsupported schedule {schedule!r}")
        self.schedule = schedule

    def get_param_value(self, timestep):
        assert timestep >= 0
        if self.type == "constant":
            return self.schedule
        knots = self.schedule
        if timestep <= knots[0][0]:
            return knots[0][1]
        if timestep >= knots[-1][0]:
            return knots[-1][1]
        for (t0, v0), (t1, v1) in zip(knots[:-1], knots[1:]):
            if t0 <= timestep < t1:
                w = float(timestep - t0) / (t1 - t0)
                return v0 + w * (v1 - v0)
        raise AssertionError("unreachable")
