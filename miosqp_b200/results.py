"""Result record of MIOSQP.solve(), same six fields as /root/reference/miosqp/results.py:1-12."""


class Results(object):
    __slots__ = ("x", "upper_glob", "run_time", "status", "osqp_solve_time", "osqp_iter_avg")

    def __init__(self, x, upper_glob, run_time, status, osqp_solve_time, osqp_iter_avg):
        self.x, self.upper_glob, self.run_time = x, upper_glob, run_time
        self.status, self.osqp_solve_time, self.osqp_iter_avg = status, osqp_solve_time, osqp_iter_avg

    def __repr__(self):
        return "Results(status=%r, upper_glob=%r, run_time=%.4g, osqp_solve_time=%.4g, osqp_iter_avg=%.4g)" % (
            self.status, self.upper_glob, self.run_time, self.osqp_solve_time, self.osqp_iter_avg)
