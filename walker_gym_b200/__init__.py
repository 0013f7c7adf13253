"""walker_gym_b200 -- B200-native batched physics step for walker-gym's PhysicsEnv.

(The distribution is called ``walker-gym_b200``; the importable package name
replaces the hyphen, which Python identifiers cannot contain.)

Drop-in surface (same names and meaning as gym/optimized_env.py,
gym/optimized_walker.py and gym/optimized_engine.py of the reference):
``make_env``, ``PhysicsEnv``, ``Environment``, ``Creature``, ``Muscle``,
``Skeleton``, ``Point``, ``DingPoint``, ``Config``, ``create_balance_creature``,
``create_box_creature``; plus ``BatchedPhysicsEnv`` for millions of envs.
"""
from .engine import Config, DingPoint, Point
from .walker import (BODIES, Creature, Muscle, Skeleton, create_balance_creature, create_box_creature,
                     make_creature)
from .batched import BatchedPhysicsEnv, HostStepPipeline, StepGraph, creature_from_id, make_params
from .actions import CPGActions, ScriptedActions
from .env import Environment, PhysicsEnv, make_env

__all__ = ["Config", "Point", "DingPoint", "Creature", "Muscle", "Skeleton", "BODIES", "make_creature",
           "create_balance_creature", "create_box_creature", "BatchedPhysicsEnv", "HostStepPipeline", "StepGraph", "creature_from_id",
           "make_params", "PhysicsEnv", "Environment", "make_env", "ScriptedActions", "CPGActions"]
