"""Backend GA behind `POST /api/simulation/start` — B200 drop-in.

Public surface kept from uav-backend/app/routers/GA_enhanced.py (GAE) so that
`uav-backend/app/routers/simulation.py` (SIMR:21-24) works after an import swap:

    GeneticAlgorithm(population_size, generations, mutation_rate).run(source, destination, environment)
    Environment(x_bounds, y_bounds, z_bounds, num_random_obstacles=30)      .obstacles -> list of voxels
    UAVPathPlanner(source, destination, environment)                        .find_path(), .dodge_obstacle()

Division of labour.  Upstream scores the WHOLE population inside `selection()` once per child pair
(GAE:28-34): P/2 * P evaluations of `fitness_function` per generation, and rebuilds the P roulette weights
every time.  Here a generation is scored by ONE `tf_fitness_gae` launch over the ragged population and its
roulette table (cumulative weights s / total, accumulated left to right like random.choices does) by ONE
`tf_roulette_cumweights` launch (`_PopulationScorer`); each `selection()` is then a bisect on that table with
the two `random.random()` draws upstream would take.  Splice and the dodge mutation are O(len) list edits
that stay on the host and pull from `random` / `np.random` in upstream's order — their draw counts depend on
the parents just selected, so the uniforms of a whole generation cannot be drawn ahead and the device-side
`tf_select_roulette` (same bisect, one thread per draw) serves the array API, not this loop.  A run under the
same seeds returns upstream's path (tests/test_gpu_dropin_gae.py, golden recorded from the unmodified file).

Upstream behaviour that is reproduced on purpose
  * roulette weights are score / total with both negative, so WORSE paths are favoured (GAE:30-33);
  * `find_path` appends the point it is standing on again at the start of every straight run (GAE:166-179);
    the whole initial population is one `tf_find_path_gae` launch (the reference's greedy walker on the device);
  * fitness is the integer  -len - #{i : path[i-1] != path[i+1]}  (GAE:20-26).
Deliberate differences
  * `mutate` dodges inside the environment given to `run()`; upstream reads a module-level global that
    only exists because the file executes a demo at import (GAE:47, GAE:181-194) — not reproduced;
  * no matplotlib plotting.
"""
import itertools
import random

import numpy as np
import torch

from . import batched as B

_AXIS_STEPS = ((1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1))


class _PopulationScorer:
    """Ragged list-of-paths -> int fitness per path, one kernel launch (tf_fitness_gae)."""

    @staticmethod
    def scores(paths):
        count = len(paths)
        width = max([len(p) for p in paths] + [1])
        packed = np.zeros((count, width, 4), np.float32)
        lengths = np.fromiter((len(p) for p in paths), np.int32, count)
        for row, path in zip(packed, paths):
            if path:
                row[:len(path), :3] = np.asarray(path, np.float32).reshape(len(path), -1)[:, :3]
        dev = B.fitness_gae(torch.from_numpy(packed).cuda(), torch.from_numpy(lengths).cuda())
        return dev.cpu().tolist()

    @staticmethod
    def scores_and_roulette(paths):
        """(scores, cumulative roulette weights) of a generation: tf_fitness_gae + tf_roulette_cumweights."""
        count = len(paths)
        width = max([len(p) for p in paths] + [1])
        packed = np.zeros((count, width, 4), np.float32)
        lengths = np.fromiter((len(p) for p in paths), np.int32, count)
        for row, path in zip(packed, paths):
            if path:
                row[:len(path), :3] = np.asarray(path, np.float32).reshape(len(path), -1)[:, :3]
        dev = B.fitness_gae(torch.from_numpy(packed).cuda(), torch.from_numpy(lengths).cuda())
        cum = B.roulette_cumweights(dev)
        return dev.cpu().tolist(), cum.cpu().tolist()


class GeneticAlgorithm:
    def __init__(self, population_size, generations, mutation_rate):
        self.population_size, self.generations, self.mutation_rate = population_size, generations, mutation_rate
        self.population = []
        self._world = None          # environment of the current run (for the dodge mutation)
        self._generation_scores = None
        self._generation_cum = None

    # -- upstream's method names -------------------------------------------------------------
    def initialize_population(self, source, destination, environment):
        """GAE:13-18: population_size walks of UAVPathPlanner.find_path.  With integer endpoints and bounds they run
        in one tf_find_path_gae launch (same paths, same consumption of `random`: the dodges read a natively drawn
        shuffle tape); anything else takes the host planner below."""
        def whole(values):
            return all(isinstance(v, (int, np.integer)) and not isinstance(v, bool) for v in values)
        limits = (environment.x_bounds, environment.y_bounds, environment.z_bounds)
        if self.population_size > 0 and whole(source) and whole(destination) and all(whole(b) for b in limits) \
                and all(whole(o) for o in environment.obstacles):
            first = tuple(source)
            for trail in B.find_path_gae(environment.obstacles, limits, source, destination, self.population_size):
                trail[0] = first                                  # upstream keeps the caller's tuple first
                self.population.append(trail)
            return
        self.population.extend(UAVPathPlanner(source, destination, environment).find_path()
                               for _ in range(self.population_size))

    def fitness_function(self, path):
        return _PopulationScorer.scores([path])[0]

    def selection(self):
        cum = self._generation_cum
        if cum is None:
            _, cum = _PopulationScorer.scores_and_roulette(self.population)
        # random.choices(population, [s / total ...], k=2) == the same call with the accumulated weights (GAE:30-33)
        return random.choices(self.population, cum_weights=cum, k=2)

    def crossover(self, parent1, parent2):
        cut = random.randint(1, min(len(parent1), len(parent2)) - 1)
        return parent1[:cut] + parent2[cut:], parent2[:cut] + parent1[cut:]

    def mutate(self, path):
        if len(path) <= 2 or random.random() >= self.mutation_rate:
            return
        where = random.randint(1, len(path) - 2)
        anchor = path[where - 1]
        path[where] = UAVPathPlanner(anchor, path[-1], self._world).dodge_obstacle(anchor, path[-1])

    def evolve(self):
        # 2 launches, reused by every selection() of the generation
        self._generation_scores, self._generation_cum = _PopulationScorer.scores_and_roulette(self.population)
        offspring = []
        for _ in range(self.population_size // 2):
            pair = self.crossover(*self.selection())
            for child in pair:
                self.mutate(child)
            offspring += pair
        self.population, self._generation_scores, self._generation_cum = offspring, None, None

    def run(self, source, destination, environment):
        self._world = environment
        self.initialize_population(source, destination, environment)
        champion = None
        for number in range(1, self.generations + 1):
            self.evolve()
            scores = _PopulationScorer.scores(self.population)
            top = scores.index(max(scores))             # first maximum, like max(population, key=...)
            champion = self.population[top]
            print(f"Generation {number}, Best Fitness: {scores[top]}")
        return champion


class Environment:
    """Voxel-set world: two cuboids around the centre plus random single voxels (np.random draws in
    upstream's order: x, y, z per voxel)."""

    def __init__(self, x_bounds, y_bounds, z_bounds, num_random_obstacles=30):
        self.x_bounds, self.y_bounds, self.z_bounds = x_bounds, y_bounds, z_bounds
        self.num_random_obstacles = num_random_obstacles
        self.obstacles = self.generate_obstacles()

    def generate_obstacles(self):
        cx, cy, cz = ((b[0] + b[1]) // 2 for b in (self.x_bounds, self.y_bounds, self.z_bounds))
        voxels = [(cx + i, cy + j, cz + k) for i, j, k in itertools.product(range(-3, 4), range(-3, 4), range(-1, 2))]
        voxels += [(cx + 5 + i, cy - 5 + j, cz + k) for i, j, k in itertools.product(range(-2, 3), repeat=3)]
        for _ in range(self.num_random_obstacles):
            voxels.append(tuple(np.random.randint(*b) for b in (self.x_bounds, self.y_bounds, self.z_bounds)))
        return list(set(voxels))


class UAVPathPlanner:
    """Straight-line walker with a random one-step dodge; seeds the population and drives `mutate`."""

    def __init__(self, source, destination, environment):
        self.source, self.destination, self.environment = source, destination, environment
        self.path = [source]

    def is_within_bounds(self, point):
        limits = (self.environment.x_bounds, self.environment.y_bounds, self.environment.z_bounds)
        return all(lo <= v <= hi for v, (lo, hi) in zip(point, limits))

    def is_collision(self, point):
        return point in self.environment.obstacles

    def compute_straight_line(self, start, end):
        span = [e - s for s, e in zip(start, end)]
        steps = max(abs(d) for d in span)
        return [tuple(int(round(s + t * d)) for s, d in zip(start, span)) for t in np.linspace(0, 1, steps + 1)]

    def dodge_obstacle(self, current, target):
        order = list(_AXIS_STEPS)
        random.shuffle(order)                       # same draw as upstream: shuffling a 6-element list
        for step in order:
            candidate = tuple(c + d for c, d in zip(current, step))
            if self.is_within_bounds(candidate) and not self.is_collision(candidate):
                return candidate
        return current

    def find_path(self):
        here = self.source
        while here != self.destination:
            for point in self.compute_straight_line(here, self.destination):
                if self.is_collision(point):
                    here = self.dodge_obstacle(here, self.destination)
                    break
                self.path.append(point)
                here = point
                if here == self.destination:
                    break
        return self.path
