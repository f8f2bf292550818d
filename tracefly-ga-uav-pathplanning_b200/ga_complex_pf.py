"""Drop-in for UAV_Navigation/path_planning_GA/"GA_complex _PF.py" (PF) — same class names,
constructor signatures, methods and quirks — with the fitness work on the B200:

    ComplexEnvironment.is_obstacle / distance_to_obstacles  PF:28-59  -> tf_point_query_f64
    Path.calculate_fitness(environment, wind_data=None)     PF:144-170 -> tf_fitness_f64 (PF rules)
    GeneticAlgorithm.evolve / best_path                     PF:210-223 -> one tf_fitness_f64 sweep of the
                                                                           whole (ragged) population per generation

Reference quirks that are reproduced on purpose:
  * evolve() appends the two children and then truncates to population_size (PF:216-218), so the
    children are discarded and the population never changes; only the np.random draws of
    crossover / mutate are consumed.  We consume exactly the same draws and launch the same G
    fitness sweeps.
  * select_parents() sorts the population in place, stably, by the PREVIOUS sweep's fitness
    (all zeros on the first call, PF:192-194).
  * no distance == 0 guard (PF:163): an all-identical path scores inf.
The random obstacle generator (PF:17-26) and the greedy PathPlanner (PF:89-132; sequential,
initial population only — SURVEY §8f-1) stay on the host and draw from `random` / `np.random`
in the reference's order; their obstacle tests are device point queries.
"""
import time  # noqa: F401  (PF:1)
from random import choice

import numpy as np
import torch

from . import _lib as L
from . import batched as B


class ComplexEnvironment:
    def __init__(self, x_bounds, y_bounds, z_bounds, num_obstacles):
        self.x_bounds = x_bounds
        self.y_bounds = y_bounds
        self.z_bounds = z_bounds
        self.num_obstacles = num_obstacles
        self.obstacles = self.generate_complex_obstacles()
        self._dev = None
        self._dev_key = None

    def generate_complex_obstacles(self):  # PF:17-26 (same draw order: choice, randint x3, randint)
        obstacles = []
        for _ in range(self.num_obstacles):
            shape = choice(["cube", "sphere", "cylinder"])
            x = np.random.randint(*self.x_bounds)
            y = np.random.randint(*self.y_bounds)
            z = np.random.randint(*self.z_bounds)
            size = np.random.randint(1, 5)
            obstacles.append({"shape": shape, "position": (x, y, z), "size": size})
        return obstacles

    def device_env(self):
        """The tf_env for the current obstacle list (rebuilt when the list is edited)."""
        key = repr(self.obstacles)
        if self._dev is None or key != self._dev_key:
            self._dev = B.Environment.from_pf_obstacles(self.obstacles)
            self._dev_key = key
        return self._dev

    def is_obstacle(self, point):  # PF:28-44
        if not self.obstacles:
            return False
        coll, _ = B.point_query_f64([tuple(point)], self.device_env())
        return bool(coll[0])

    def distance_to_obstacles(self, point):  # PF:46-59
        if not self.obstacles:
            return float("inf")
        _, dn = B.point_query_f64([tuple(point)], self.device_env())
        return float(dn[0])

    def plot_environment(self, ax):  # PF:61-85 — plotting only, needs matplotlib's 3-D toolkit
        from mpl_toolkits.mplot3d.art3d import Poly3DCollection
        for obs in self.obstacles:
            shape, position, size = obs["shape"], obs["position"], obs["size"]
            if shape == "cube":
                x, y, z = position
                r = [-size, size]
                verts = [[(x + dx, y + dy, z + dz) for dx in r for dy in r for dz in r]]
                ax.add_collection3d(Poly3DCollection(verts, alpha=0.5, color="gray"))
            elif shape == "sphere":
                u, v = np.mgrid[0:2 * np.pi:20j, 0:np.pi:10j]
                ax.plot_surface(size * np.cos(u) * np.sin(v) + position[0], size * np.sin(u) * np.sin(v) + position[1],
                                size * np.cos(v) + position[2], color="gray", alpha=0.5)
            elif shape == "cylinder":
                x = np.linspace(-size, size, 50) + position[0]
                z = np.linspace(0, 2 * size, 50) + position[2]
                X, Z = np.meshgrid(x, z)
                ax.plot_surface(X, np.sqrt(size ** 2 - X ** 2) + position[1], Z, color="gray", alpha=0.5)


class PathPlanner:
    """PF:89-132: greedy sign-step walk with a one-move dodge.  Host-side and sequential; used only
    to seed the initial population."""

    MOVES = {"up": (0, 0, 1), "down": (0, 0, -1), "left": (-1, 0, 0), "right": (1, 0, 0),
             "forward": (0, 1, 0), "backward": (0, -1, 0)}

    def __init__(self, source, destination, environment):
        self.source = source
        self.destination = destination
        self.environment = environment

    def find_path(self, dodge_strategy, height_variation=0):
        current = self.source
        path = [current]
        while current != self.destination:
            next_point = self.find_next_step(current, self.destination, height_variation)
            if self.environment.is_obstacle(next_point):
                current = self.dodge_obstacle(current, dodge_strategy)
            else:
                current = next_point
            if current in path:
                break  # PF:108-109
            path.append(current)
        return path

    def find_next_step(self, current, target, height_variation):  # PF:113-119
        direction = (np.sign(target[0] - current[0]), np.sign(target[1] - current[1]),
                     np.sign(target[2] - current[2] + height_variation))
        return (current[0] + direction[0], current[1] + direction[1], current[2] + direction[2])

    def dodge_obstacle(self, current, dodge_strategy):  # PF:121-132
        move = self.MOVES[dodge_strategy]
        next_point = (current[0] + move[0], current[1] + move[1], current[2] + move[2])
        return next_point if not self.environment.is_obstacle(next_point) else current


def _device_population(paths):
    P = len(paths)
    Wmax = max(1, max(len(p) for p in paths))
    host = np.zeros((P, Wmax, 4), np.float64)
    lens = np.zeros(P, np.int32)
    for i, p in enumerate(paths):
        lens[i] = len(p)
        if len(p):
            host[i, :len(p), :3] = np.asarray(p, np.float64).reshape(len(p), -1)[:, :3]
    return torch.from_numpy(host).cuda(), torch.from_numpy(lens).cuda()


class Path:
    def __init__(self, waypoints):
        self.waypoints = waypoints
        self.fitness = 0

    def is_turning(self, i):  # PF:141-142
        return self.waypoints[i][0] != self.waypoints[i - 1][0] or self.waypoints[i][1] != self.waypoints[i - 1][1]

    def calculate_fitness(self, environment, wind_data=None):  # PF:144-170 (wind_data is unused upstream too)
        pop, lens = _device_population([self.waypoints])
        fit, _ = B.fitness_batch_f64(pop, lens, environment.device_env(), L.default_params(dist_zero_guard=0))
        self.fitness = float(fit[0].item())


class GeneticAlgorithm:
    def __init__(self, population_size, mutation_rate, generations, environment, source, destination):
        self.population_size = population_size
        self.mutation_rate = mutation_rate
        self.generations = generations
        self.environment = environment
        self.source = source
        self.destination = destination
        self.population = self.initialize_population()

    def initialize_population(self):  # PF:184-190 (one randint(-2, 2) per individual, in order)
        population = []
        for _ in range(self.population_size):
            planner = PathPlanner(self.source, self.destination, self.environment)
            path = planner.find_path(dodge_strategy="up", height_variation=np.random.randint(-2, 2))
            population.append(Path(path))
        return population

    def select_parents(self):  # PF:192-194
        self.population.sort(key=lambda x: x.fitness, reverse=True)
        return self.population[:2]

    def crossover(self, parent1, parent2):  # PF:196-200
        crossover_point = np.random.randint(1, len(parent1.waypoints) - 1)
        child1 = parent1.waypoints[:crossover_point] + parent2.waypoints[crossover_point:]
        child2 = parent2.waypoints[:crossover_point] + parent1.waypoints[crossover_point:]
        return Path(child1), Path(child2)

    def mutate(self, child):  # PF:202-208
        if np.random.rand() < self.mutation_rate:
            mutation_point = np.random.randint(0, len(child.waypoints))
            w = child.waypoints[mutation_point]
            child.waypoints[mutation_point] = (w[0] + np.random.choice([-1, 1]), w[1] + np.random.choice([-1, 1]),
                                               w[2] + np.random.choice([-1, 1]))
        return child

    def _sweep(self):
        """calculate_fitness for the whole (ragged) population: one tf_fitness_f64 launch."""
        pop, lens = _device_population([p.waypoints for p in self.population])
        fit, _ = B.fitness_batch_f64(pop, lens, self.environment.device_env(), L.default_params(dist_zero_guard=0))
        for p, f in zip(self.population, fit.cpu().tolist()):
            p.fitness = f

    def evolve(self):  # PF:210-220
        for _generation in range(self.generations):
            parents = self.select_parents()
            child1, child2 = self.crossover(parents[0], parents[1])
            child1 = self.mutate(child1)
            child2 = self.mutate(child2)
            self.population.append(child1)
            self.population.append(child2)
            self.population = self.population[:self.population_size]   # PF:218 drops both children
            self._sweep()

    def best_path(self):  # PF:222-223
        return max(self.population, key=lambda x: x.fitness)
