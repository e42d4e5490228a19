"""Parallel runtime: arenas, placement, symmetric memory, PS/worker roles, engines, Trainer."""
from .arena import ArenaLayout, ModelBinder
from .placement import Placement

__all__ = ["ArenaLayout", "ModelBinder", "Placement", "Trainer"]


def __getattr__(name):
    if name == "Trainer":
        from .trainer import Trainer
        return Trainer
    raise AttributeError(name)
