"""Shaped-obstacle GA — B200 drop-in for UAV_Navigation/path_planning_GA/"GA_complex _PF.py" (PF).

Same public classes and signatures (`ComplexEnvironment`, `PathPlanner`, `Path`, `GeneticAlgorithm`), with
the obstacle arithmetic on the GPU:

    ComplexEnvironment.is_obstacle / distance_to_obstacles   PF:28-59    tf_point_query_f64
    Path.calculate_fitness(environment, wind_data=None)      PF:144-170  tf_fitness_f64 (PF rules)
    GeneticAlgorithm.evolve / best_path                      PF:210-223  one tf_fitness_f64 sweep of the whole
                                                                         ragged population per generation

Upstream behaviour that is reproduced on purpose
  * `evolve` appends both children and then cuts the list back to `population_size` (PF:216-218): the
    children never survive, the population never changes, and only the np.random draws of crossover /
    mutate are consumed — we consume the same draws and launch the same G sweeps;
  * `select_parents` sorts in place, stably, on the PREVIOUS sweep's fitness (zeros the first time);
  * no zero-distance guard: an all-identical path scores inf (PF:163).
    PathPlanner.find_path                                    PF:93-107   tf_find_path_pf (one thread per path);
    GeneticAlgorithm.initialize_population                   PF:181-188  ONE launch for the whole population
The random obstacle generator (PF:17-26) is host code that draws from `random` / `np.random` in upstream's
order.  `find_next_step` / `dodge_obstacle` stay available as single-step host methods (their obstacle test
is a device point query).  Plotting helpers are not provided.
"""
import random

import numpy as np
import torch

from . import _lib as L
from . import batched as B

_SHAPES = ("cube", "sphere", "cylinder")
_DODGES = dict(up=(0, 0, 1), down=(0, 0, -1), left=(-1, 0, 0), right=(1, 0, 0), forward=(0, 1, 0), backward=(0, -1, 0))
_PF_PARAMS = dict(dist_zero_guard=0)     # PF:163 divides by the distance unguarded


def _upload(paths):
    """list of waypoint lists -> (f64[P, Wmax, 4] cuda, i32[P] cuda)"""
    lens = np.fromiter((len(p) for p in paths), np.int32, len(paths))
    block = np.zeros((len(paths), max(int(lens.max(initial=0)), 1), 4))
    for dst, p in zip(block, paths):
        if p:
            dst[:len(p), :3] = np.asarray(p, np.float64).reshape(len(p), -1)[:, :3]
    return torch.from_numpy(block).cuda(), torch.from_numpy(lens).cuda()


class ComplexEnvironment:
    def __init__(self, x_bounds, y_bounds, z_bounds, num_obstacles):
        self.x_bounds, self.y_bounds, self.z_bounds = x_bounds, y_bounds, z_bounds
        self.num_obstacles = num_obstacles
        self.obstacles = self.generate_complex_obstacles()
        self._device = (None, None)

    def generate_complex_obstacles(self):
        # draw order per obstacle, as upstream: random.choice(shape), randint x, y, z, randint size in [1, 4]
        made = []
        for _ in range(self.num_obstacles):
            shape = random.choice(list(_SHAPES))
            centre = tuple(np.random.randint(*b) for b in (self.x_bounds, self.y_bounds, self.z_bounds))
            made.append({"shape": shape, "position": centre, "size": np.random.randint(1, 5)})
        return made

    def device_env(self):
        """tf_env of the current obstacle list (rebuilt if the list was edited)."""
        stamp = repr(self.obstacles)
        if self._device[0] != stamp:
            self._device = (stamp, B.Environment.from_pf_obstacles(self.obstacles))
        return self._device[1]

    def _query(self, point):
        return B.point_query_f64([tuple(point)], self.device_env())

    def is_obstacle(self, point):
        return bool(self.obstacles) and bool(self._query(point)[0][0])

    def distance_to_obstacles(self, point):
        return float(self._query(point)[1][0]) if self.obstacles else float("inf")


class PathPlanner:
    """Greedy sign-step walk towards the destination with a one-move dodge (host, sequential)."""

    def __init__(self, source, destination, environment):
        self.source, self.destination, self.environment = source, destination, environment

    def find_next_step(self, current, target, height_variation):
        lift = (0, 0, height_variation)
        return tuple(c + np.sign(t - c + b) for c, t, b in zip(current, target, lift))

    def dodge_obstacle(self, current, dodge_strategy):
        shifted = tuple(c + d for c, d in zip(current, _DODGES[dodge_strategy]))
        return current if self.environment.is_obstacle(shifted) else shifted

    def find_paths(self, dodge_strategy, height_variations):
        """`find_path` for every entry of `height_variations`: one tf_find_path_pf launch."""
        if dodge_strategy not in _DODGES:
            raise KeyError(dodge_strategy)                      # upstream: moves[dodge_strategy]
        paths, lens = B.find_path_pf(self.environment.device_env(), self.source, self.destination,
                                     [int(v) for v in height_variations], dodge_strategy)
        rows, counts = paths.cpu().numpy(), lens.cpu().tolist()
        whole = all(float(v).is_integer() for v in tuple(self.source) + tuple(self.destination))
        out = []
        for row, count in zip(rows, counts):
            pts = row[:count, :3]
            trail = [tuple(int(v) for v in q) for q in pts] if whole else [tuple(float(v) for v in q) for q in pts]
            trail[0] = self.source                              # upstream keeps the caller's tuple first
            out.append(trail)
        return out

    def find_path(self, dodge_strategy, height_variation=0):
        return self.find_paths(dodge_strategy, [height_variation])[0]


class Path:
    def __init__(self, waypoints):
        self.waypoints = waypoints
        self.fitness = 0

    def is_turning(self, i):
        (x1, y1, *_), (x0, y0, *_) = self.waypoints[i], self.waypoints[i - 1]
        return x1 != x0 or y1 != y0

    def calculate_fitness(self, environment, wind_data=None):   # wind_data is unused upstream as well
        pop, lens = _upload([self.waypoints])
        score, _ = B.fitness_batch_f64(pop, lens, environment.device_env(), L.default_params(**_PF_PARAMS))
        self.fitness = float(score[0].item())


class GeneticAlgorithm:
    def __init__(self, population_size, mutation_rate, generations, environment, source, destination):
        self.population_size, self.mutation_rate, self.generations = population_size, mutation_rate, generations
        self.environment, self.source, self.destination = environment, source, destination
        self.population = self.initialize_population()

    def initialize_population(self):
        # one np.random.randint(-2, 2) per individual, drawn in upstream's order; the walks themselves consume
        # no random numbers, so all of them run in one launch
        lifts = [np.random.randint(-2, 2) for _ in range(self.population_size)]
        planner = PathPlanner(self.source, self.destination, self.environment)
        return [Path(trail) for trail in planner.find_paths("up", lifts)]

    def select_parents(self):
        self.population.sort(key=lambda p: p.fitness, reverse=True)
        return self.population[:2]

    def crossover(self, parent1, parent2):
        cut = np.random.randint(1, len(parent1.waypoints) - 1)
        a, b = parent1.waypoints, parent2.waypoints
        return Path(a[:cut] + b[cut:]), Path(b[:cut] + a[cut:])

    def mutate(self, child):
        if np.random.rand() < self.mutation_rate:
            at = np.random.randint(0, len(child.waypoints))
            child.waypoints[at] = tuple(v + np.random.choice([-1, 1]) for v in child.waypoints[at][:3])
        return child

    def _score_population(self):
        """`calculate_fitness` of every individual: one tf_fitness_f64 launch over the ragged population."""
        pop, lens = _upload([p.waypoints for p in self.population])
        scores, _ = B.fitness_batch_f64(pop, lens, self.environment.device_env(), L.default_params(**_PF_PARAMS))
        for individual, value in zip(self.population, scores.cpu().tolist()):
            individual.fitness = value

    def evolve(self):
        for _ in range(self.generations):
            first, second = self.select_parents()
            self.population += [self.mutate(c) for c in self.crossover(first, second)]
            del self.population[self.population_size:]          # upstream's truncation drops both children
            self._score_population()

    def best_path(self):
        return max(self.population, key=lambda p: p.fitness)
