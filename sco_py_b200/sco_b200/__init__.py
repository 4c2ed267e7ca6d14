"""The B200 backend package: same triple (prob.Prob, solver.Solver, variable.Variable + a scalar
variable type) and method surfaces as sco_py.sco_osqp / sco_py.sco_gurobi, sharing sco_py_b200.expr."""
