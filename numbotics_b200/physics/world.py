"""Obstacle / chain registry the collision path reads.

Mirror of the registry half of reference numbotics/physics/world.py:18-57,
208-260 (`WORLD_INSTANCES`, `get_world`, `register`, `get_object`).  The reference
`World` also owns a PyBullet client, stepping, pooling and a visualiser; none of
that is on the data-parallel path (out of scope), so here a
`World` is only the set of bodies whose geometry and pose `Arm.in_collision`
consumes, plus a monotonically increasing `revision` that `Arm` uses to know when
the device scene descriptor is stale.
"""
import weakref
from typing import Any

WORLD_INSTANCES: dict = {}
SELECTED_WORLD = None


def get_world(name=None):
    global SELECTED_WORLD
    if name is None:
        name = SELECTED_WORLD
    if name is None and len(WORLD_INSTANCES) == 0:
        name = "world_0"
    if name is None:
        raise ValueError("There should always be a selected World instance if a World exists")
    if name not in WORLD_INSTANCES:
        World(name=name)
    SELECTED_WORLD = name
    return WORLD_INSTANCES[name]


class World:

    def __init__(self, name=None, visualize: bool = False):
        global SELECTED_WORLD
        if visualize:
            raise NotImplementedError("visualisation is outside the batched kinematics/collision path")
        self._name = f"world_{len(WORLD_INSTANCES)}" if name is None else name
        if name in WORLD_INSTANCES:
            raise ValueError(f"World with name {name} already exists")
        WORLD_INSTANCES[self._name] = self
        SELECTED_WORLD = self._name
        self._static_objects = weakref.WeakValueDictionary()
        self._dynamic_objects = weakref.WeakValueDictionary()
        self._entity_map = weakref.WeakValueDictionary()
        self._next_id = 0
        self.revision = 0
        self.gravity = None

    @property
    def name(self):
        return self._name

    def _new_entity_id(self):
        i = self._next_id
        self._next_id += 1
        return i

    def touch(self):
        """Bodies call this whenever geometry or pose changes."""
        self.revision += 1

    def register(self, body):
        from .object import PhysicsObject
        from .chain import Chain
        if not isinstance(body, (PhysicsObject, Chain)):
            raise ValueError(f"Unknown body type: {type(body)}")
        if body._static:
            self._static_objects[body.name] = body
        else:
            self._dynamic_objects[body.name] = body
        self._entity_map[body._entity_id] = body
        self.touch()

    def unregister(self, body):
        for d in (self._static_objects, self._dynamic_objects):
            if body.name in d:
                del d[body.name]
        self._entity_map.pop(body._entity_id, None)
        self.touch()

    def objects(self):
        """Static then dynamic bodies, the iteration order of reference robots/arm.py:228,563."""
        return list(self._static_objects.values()) + list(self._dynamic_objects.values())

    def get_object(self, name: str, default: Any = None):
        from .chain import Chain
        if name in self._static_objects:
            return self._static_objects[name]
        if name in self._dynamic_objects:
            return self._dynamic_objects[name]
        for obj in list(self._dynamic_objects.values()):
            if isinstance(obj, Chain):
                for link in obj._links:
                    if link.name == name:
                        return link
        return default

    def clear(self):
        self._static_objects = weakref.WeakValueDictionary()
        self._dynamic_objects = weakref.WeakValueDictionary()
        self._entity_map = weakref.WeakValueDictionary()
        self.touch()

    def step(self, sleep: bool = False):
        """No dynamics on this path; kept so reference scripts that call it still run."""
        return None


def _reset_worlds():
    """Test helper: forget every world."""
    global SELECTED_WORLD
    WORLD_INSTANCES.clear()
    SELECTED_WORLD = None
