"""The reference's *package* lineage (``gym/optimized_walker/``) on the B200 library.

Same names as ``gym/optimized_walker/{core,env,walker}.py``: ``Config``,
``Point``, ``DingPoint``, ``to_data``, ``Environment``, ``OptimizedEnvironment``,
``Muscle``, ``Skeleton``, ``Creature``, ``Brain`` and the body builders
``test, leg2, box, balance1, balance2, balance3, humanb, insect``.
``Environment(..., num_envs=E)`` steps E independent copies per call.
Rendering (pygame) is out of scope.
"""
from .core import Config, DingPoint, Point, to_data
from .env import Environment, OptimizedEnvironment
from .walker import (Brain, Creature, Muscle, Skeleton, balance1, balance2, balance3, box, humanb, insect, leg2,
                     test)

__all__ = ["Config", "Point", "DingPoint", "to_data", "Environment", "OptimizedEnvironment", "Muscle", "Skeleton",
           "Creature", "Brain", "test", "leg2", "box", "balance1", "balance2", "balance3", "humanb", "insect"]
