"""Split divergence operators (src/discrete_operators.py)."""


def divergence(cs_grid, simulation):
    """src/discrete_operators.py:18-101, operator by operator on the device."""
    simulation.dev.call("pycs_divergence")


def F_operator(cs_grid, simulation):
    """src/discrete_operators.py:109-120."""
    simulation.dev.call("pycs_F_operator")


def G_operator(cs_grid, simulation):
    """src/discrete_operators.py:128-138."""
    simulation.dev.call("pycs_G_operator")
