"""Host-side road description: lanes, the road network graph and its flat device table.

Mirrors the constructor surface of the reference's ``ttrl_env.road.lane`` / ``ttrl_env.road.road``
(``StraightLane`` lane.py:159-194, ``SineLane`` :236-266, ``CircularLane`` :311-339,
``RoadNetwork.add_lane`` road.py:27-39, ``straight_road_network`` road.py:291-321,
``shortest_path`` road.py:159-188) so scenes are written the same way, but holds no per-step logic:
every geometric query on the step path runs on the device from the table built by
:meth:`RoadNetwork.to_table`.  Derived constants (heading, length, unit direction) are computed with the
same float64 numpy expressions as the reference so the uploaded doubles are bit-identical to the
attributes the reference would hold.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import abi

LaneIndex = Tuple[str, str, Optional[int]]


class LineType:
    NONE = 0
    STRIPED = 1
    CONTINUOUS = 2
    CONTINUOUS_LINE = 3


def wrap_to_pi(x: float) -> float:
    return ((x + np.pi) % (2 * np.pi)) - np.pi  # utils.py:57-58


class AbstractLane:
    DEFAULT_WIDTH: float = 4
    VEHICLE_LENGTH: float = 5
    length: float = 0.0

    def _record(self) -> abi.Lane:  # pragma: no cover - abstract
        raise NotImplementedError

    # Scene-construction helpers (host-driven resets place vehicles like RoadObject.__init__, objects.py:45-50, which
    # looks up the closest lane of the new vehicle); never on the step path.
    def local_coordinates(self, position) -> Tuple[float, float]:  # pragma: no cover - abstract
        raise NotImplementedError

    def distance_with_heading(self, position, heading: float, heading_weight: float = 1.0) -> float:  # lane.py:132-147
        s, r = self.local_coordinates(position)
        angle = np.abs(wrap_to_pi(heading - self.heading_at(s)))
        return abs(r) + max(s - self.length, 0) + max(0 - s, 0) + heading_weight * angle


class StraightLane(AbstractLane):
    def __init__(self, start, end, width: float = AbstractLane.DEFAULT_WIDTH, line_types=None,
                 forbidden: bool = False, speed_limit: float = 20, priority: int = 0) -> None:
        self.start = np.array(start, dtype=np.float64)
        self.end = np.array(end, dtype=np.float64)
        self.width = width
        delta = self.end - self.start
        self.heading = float(np.arctan2(delta[1], delta[0]))
        self.length = float(np.linalg.norm(delta))
        self.direction = delta / self.length
        self.direction_lateral = np.array([-self.direction[1], self.direction[0]])
        self.line_types = line_types or [LineType.STRIPED, LineType.STRIPED]
        self.forbidden = forbidden
        self.priority = priority
        self.speed_limit = speed_limit

    def _record(self) -> abi.Lane:
        rec = abi.Lane()
        rec.kind = abi.LANE_STRAIGHT
        rec.ax, rec.ay = float(self.start[0]), float(self.start[1])
        rec.dx, rec.dy = float(self.direction[0]), float(self.direction[1])
        rec.heading = self.heading
        rec.length = self.length
        return rec

    # geometry helpers used only at scene-construction time (never on the step path)
    def position(self, longitudinal: float, lateral: float) -> np.ndarray:
        return self.start + longitudinal * self.direction + lateral * self.direction_lateral

    def heading_at(self, longitudinal: float) -> float:
        return self.heading

    def local_coordinates(self, position) -> Tuple[float, float]:  # lane.py:209-213
        delta = np.asarray(position, dtype=np.float64) - self.start
        return float(np.dot(delta, self.direction)), float(np.dot(delta, self.direction_lateral))


class SineLane(StraightLane):
    def __init__(self, start, end, amplitude: float, pulsation: float, phase: float,
                 width: float = AbstractLane.DEFAULT_WIDTH, line_types=None, forbidden: bool = False,
                 speed_limit: float = 20, priority: int = 0) -> None:
        super().__init__(start, end, width, line_types, forbidden, speed_limit, priority)
        self.amplitude = amplitude
        self.pulsation = pulsation
        self.phase = phase

    def _record(self) -> abi.Lane:
        rec = super()._record()
        rec.kind = abi.LANE_SINE
        rec.amplitude, rec.pulsation, rec.phase = float(self.amplitude), float(self.pulsation), float(self.phase)
        return rec

    def position(self, longitudinal: float, lateral: float) -> np.ndarray:
        return super().position(longitudinal, lateral + self.amplitude * np.sin(self.pulsation * longitudinal + self.phase))

    def heading_at(self, longitudinal: float) -> float:
        return self.heading + np.arctan(self.amplitude * self.pulsation * np.cos(self.pulsation * longitudinal + self.phase))

    def local_coordinates(self, position) -> Tuple[float, float]:  # lane.py:282-286
        longitudinal, lateral = super().local_coordinates(position)
        return longitudinal, lateral - self.amplitude * np.sin(self.pulsation * longitudinal + self.phase)


class CircularLane(AbstractLane):
    def __init__(self, center, radius: float, start_phase: float, end_phase: float, clockwise: bool = True,
                 width: float = AbstractLane.DEFAULT_WIDTH, line_types=None, forbidden: bool = False,
                 speed_limit: float = 20, priority: int = 0) -> None:
        self.center = np.array(center, dtype=np.float64)
        self.radius = radius
        self.start_phase = start_phase
        self.end_phase = end_phase
        self.clockwise = clockwise
        self.direction = 1 if clockwise else -1
        self.width = width
        self.line_types = line_types or [LineType.STRIPED, LineType.STRIPED]
        self.forbidden = forbidden
        self.length = float(radius * (end_phase - start_phase) * self.direction)
        self.priority = priority
        self.speed_limit = speed_limit

    def _record(self) -> abi.Lane:
        rec = abi.Lane()
        rec.kind = abi.LANE_CIRCULAR
        rec.ax, rec.ay = float(self.center[0]), float(self.center[1])
        rec.radius = float(self.radius)
        rec.start_phase, rec.end_phase = float(self.start_phase), float(self.end_phase)
        rec.cdir = float(self.direction)
        rec.length = self.length
        return rec

    def position(self, longitudinal: float, lateral: float) -> np.ndarray:
        phi = self.direction * longitudinal / self.radius + self.start_phase
        return self.center + (self.radius - lateral * self.direction) * np.array([np.cos(phi), np.sin(phi)])

    def heading_at(self, longitudinal: float) -> float:
        phi = self.direction * longitudinal / self.radius + self.start_phase
        return phi + np.pi / 2 * self.direction

    def local_coordinates(self, position) -> Tuple[float, float]:  # lane.py:355-362
        delta = np.asarray(position, dtype=np.float64) - self.center
        phi = np.arctan2(delta[1], delta[0])
        phi = self.start_phase + wrap_to_pi(phi - self.start_phase)
        r = np.linalg.norm(delta)
        return float(self.direction * (phi - self.start_phase) * self.radius), float(self.direction * (self.radius - r))


@dataclass
class NetworkTable:
    """Flat, device-ready form of a :class:`RoadNetwork` (see ``ttrl_lane`` / ``ttrl_road`` in the header)."""
    lanes: C.Array
    roads: C.Array
    node_first: np.ndarray
    node_roads: np.ndarray
    node_names: List[str]
    road_keys: List[Tuple[str, str]]
    lane_keys: List[Tuple[str, str, int]]
    lane_index_of: Dict[Tuple[str, str, int], int] = field(default_factory=dict)
    road_index_of: Dict[Tuple[str, str], int] = field(default_factory=dict)

    @property
    def n_lanes(self) -> int:
        return len(self.lane_keys)

    @property
    def n_roads(self) -> int:
        return len(self.road_keys)

    @property
    def n_nodes(self) -> int:
        return len(self.node_names)

    def flat(self, index: LaneIndex) -> int:
        _from, _to, _id = index
        if _id is None:
            _id = 0
        return self.lane_index_of[(_from, _to, int(_id))]


class RoadNetwork:
    """``graph[_from][_to] -> [lanes]`` with insertion order preserved (it is semantic: closest-lane ties)."""

    def __init__(self) -> None:
        self.graph: Dict[str, Dict[str, List[AbstractLane]]] = {}

    def add_lane(self, _from: str, _to: str, lane: AbstractLane) -> None:
        self.graph.setdefault(_from, {}).setdefault(_to, []).append(lane)

    def get_lane(self, index: LaneIndex) -> AbstractLane:
        _from, _to, _id = index
        if _id is None and len(self.graph[_from][_to]) == 1:
            _id = 0
        return self.graph[_from][_to][_id]

    def lanes_list(self) -> List[AbstractLane]:
        return [lane for tos in self.graph.values() for lanes in tos.values() for lane in lanes]

    def get_closest_lane_index(self, position, heading: float) -> LaneIndex:
        """road.py:55-71: argmin of ``distance_with_heading`` in dict order (the first minimum wins)."""
        indexes, distances = [], []
        for _from, tos in self.graph.items():
            for _to, lanes in tos.items():
                for _id, lane in enumerate(lanes):
                    distances.append(lane.distance_with_heading(position, heading))
                    indexes.append((_from, _to, _id))
        return indexes[int(np.argmin(distances))]

    # ---- route planning (host side; the device receives finished route tables) --------------------
    def bfs_paths(self, start: str, goal: str):
        """Breadth-first enumeration with neighbours visited in sorted order (reference road.py:159-178)."""
        queue = [(start, [start])]
        while queue:
            node, path = queue.pop(0)
            if node not in self.graph:
                yield []
            for nxt in sorted(k for k in self.graph[node].keys() if k not in path):
                if nxt == goal:
                    yield path + [nxt]
                elif nxt in self.graph:
                    queue.append((nxt, path + [nxt]))

    def shortest_path(self, start: str, goal: str) -> List[str]:
        return next(self.bfs_paths(start, goal), [])

    def plan_route(self, lane_index: LaneIndex, destination: str) -> List[LaneIndex]:
        """``ControlledVehicle.plan_route_to`` (controller.py:71-87) as a pure function."""
        try:
            path = self.shortest_path(lane_index[1], destination)
        except KeyError:
            path = []
        if path:
            return [lane_index] + [(path[i], path[i + 1], None) for i in range(len(path) - 1)]
        return [lane_index]

    @staticmethod
    def straight_road_network(lanes: int = 4, start: float = 0, length: float = 10000, angle: float = 0,
                              speed_limit: float = 30, nodes_str: Optional[Tuple[str, str]] = None,
                              net: Optional["RoadNetwork"] = None) -> "RoadNetwork":
        net = net or RoadNetwork()
        nodes_str = nodes_str or ("0", "1")
        rotation = np.array([[np.cos(angle), np.sin(angle)], [-np.sin(angle), np.cos(angle)]])
        for lane in range(lanes):
            origin = rotation @ np.array([start, lane * AbstractLane.DEFAULT_WIDTH])
            end = rotation @ np.array([start + length, lane * AbstractLane.DEFAULT_WIDTH])
            line_types = [LineType.CONTINUOUS_LINE if lane == 0 else LineType.STRIPED,
                          LineType.CONTINUOUS_LINE if lane == lanes - 1 else LineType.NONE]
            net.add_lane(*nodes_str, StraightLane(origin, end, line_types=line_types, speed_limit=speed_limit))
        return net

    # ---- flattening ------------------------------------------------------------------------------
    def to_table(self, exit_predicate=None) -> NetworkTable:
        """Flatten in dict insertion order.  ``exit_predicate(_from, _to)`` marks exit lanes
        (IntersectionEnv: ``"il" in _from and "o" in _to``, intersection_env.py:352-353)."""
        node_names: List[str] = []

        def node(name: str) -> int:
            if name not in node_names:
                node_names.append(name)
            return node_names.index(name)

        lane_recs: List[abi.Lane] = []
        road_recs: List[abi.Road] = []
        road_keys: List[Tuple[str, str]] = []
        lane_keys: List[Tuple[str, str, int]] = []
        out_roads: Dict[int, List[int]] = {}
        for _from, tos in self.graph.items():
            for _to, lanes in tos.items():
                r = abi.Road()
                r.from_node, r.to_node = node(_from), node(_to)
                r.first_lane, r.n_lanes = len(lane_recs), len(lanes)
                ridx = len(road_recs)
                road_recs.append(r)
                road_keys.append((_from, _to))
                out_roads.setdefault(r.from_node, []).append(ridx)
                for _id, lane in enumerate(lanes):
                    rec = lane._record()
                    rec.road, rec.lane_id = ridx, _id
                    rec.priority = int(lane.priority)
                    rec.forbidden = int(bool(lane.forbidden))
                    rec.width = float(lane.width)
                    rec.speed_limit = float(lane.speed_limit)
                    rec.is_exit = int(bool(exit_predicate(_from, _to))) if exit_predicate else 0
                    lane_recs.append(rec)
                    lane_keys.append((_from, _to, _id))
        if len(lane_recs) > abi.MAX_LANES or len(road_recs) > abi.MAX_ROADS or len(node_names) > abi.MAX_NODES:
            raise ValueError("road network exceeds the device table capacity")
        node_first = np.zeros(len(node_names) + 1, dtype=np.int32)
        node_roads: List[int] = []
        for n in range(len(node_names)):
            node_first[n] = len(node_roads)
            node_roads.extend(out_roads.get(n, []))
        node_first[len(node_names)] = len(node_roads)
        table = NetworkTable(
            lanes=(abi.Lane * len(lane_recs))(*lane_recs),
            roads=(abi.Road * len(road_recs))(*road_recs),
            node_first=node_first,
            node_roads=np.asarray(node_roads + [0], dtype=np.int32),
            node_names=node_names, road_keys=road_keys, lane_keys=lane_keys,
        )
        table.lane_index_of = {k: i for i, k in enumerate(lane_keys)}
        table.road_index_of = {k: i for i, k in enumerate(road_keys)}
        return table
