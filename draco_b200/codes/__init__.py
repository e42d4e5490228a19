"""Gradient codes, adversary model and oracles."""
from .adversary import err_simulation, generate_schedule
from .cyclic import CyclicCode, search_w
from .repetition import GroupPlan, group_assign

__all__ = ["err_simulation", "generate_schedule", "CyclicCode", "search_w", "GroupPlan", "group_assign"]
