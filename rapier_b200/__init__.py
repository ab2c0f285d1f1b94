"""rapier_b200 -- B200-native step hot path behind Rapier's API."""
from . import _abi  # noqa: F401
from .sets import (ColliderBuilder, ColliderSet, FixedJointBuilder, ImpulseJointSet, RigidBodyBuilder,  # noqa: F401
                   RigidBodySet, SphericalJointBuilder, SpringJointBuilder, RopeJointBuilder)
