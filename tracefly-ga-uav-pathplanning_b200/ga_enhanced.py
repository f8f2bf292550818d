"""Drop-in for uav-backend/app/routers/GA_enhanced.py (GAE) — `GeneticAlgorithm(population_size,
generations, mutation_rate).run(source, destination, environment)`, `Environment`, `UAVPathPlanner` —
so that `uav-backend/app/routers/simulation.py` (SIMR:21-24) keeps working after an import swap.

What runs where: the reference evaluates `fitness_function` for the WHOLE population inside
`selection()` once per child pair (GAE:28-34) — P/2 * P evaluations per generation.  Here the whole
ragged population is scored by ONE `tf_fitness_gae` launch per generation; the roulette draw
(`random.choices`), the single-point splice and the dodge mutation are cheap list operations that
stay on the host and consume `random` / `np.random` exactly like the reference, so a run under the
same seeds returns the same path (tests/test_gpu_dropin_gae.py against a golden recorded from the
unmodified reference).

Reference quirks kept: the roulette weights are score/total with both negative, i.e. WORSE paths are
preferred (GAE:30-33); `find_path` re-appends the current point (GAE:166-179); integer fitness.
One deliberate fix: `mutate` uses the environment passed to `run()` — upstream reads a module-level
global that only exists because the file runs a demo at import time (GAE:47, GAE:181-194).
The file's import-time demo and matplotlib plotting are not reproduced.
"""
import random

import numpy as np
import torch

from . import batched as B


class GeneticAlgorithm:
    def __init__(self, population_size, generations, mutation_rate):
        self.population_size = population_size
        self.generations = generations
        self.mutation_rate = mutation_rate
        self.population = []
        self._environment = None
        self._scores = None

    def initialize_population(self, source, destination, environment):  # GAE:13-18
        for _ in range(self.population_size):
            planner = UAVPathPlanner(source, destination, environment)
            self.population.append(planner.find_path())

    def fitness_function(self, path):  # GAE:20-26 (single path; batches go through _score_population)
        return int(self._score([path])[0])

    @staticmethod
    def _score(paths):
        P = len(paths)
        Wmax = max(1, max(len(p) for p in paths))
        host = np.zeros((P, Wmax, 4), np.float32)
        lens = np.zeros(P, np.int32)
        for i, p in enumerate(paths):
            lens[i] = len(p)
            if len(p):
                host[i, :len(p), :3] = np.asarray(p, np.float32).reshape(len(p), -1)[:, :3]
        fit = B.fitness_gae(torch.from_numpy(host).cuda(), torch.from_numpy(lens).cuda())
        return fit.cpu().tolist()

    def selection(self):  # GAE:28-34 roulette wheel (on the population's device-computed scores)
        fitness_scores = self._scores if self._scores is not None else self._score(self.population)
        total_fitness = sum(fitness_scores)
        probabilities = [score / total_fitness for score in fitness_scores]
        return random.choices(self.population, probabilities, k=2)

    def crossover(self, parent1, parent2):  # GAE:36-41
        split_point = random.randint(1, min(len(parent1), len(parent2)) - 1)
        return parent1[:split_point] + parent2[split_point:], parent2[:split_point] + parent1[split_point:]

    def mutate(self, path):  # GAE:43-48
        if len(path) > 2 and random.random() < self.mutation_rate:
            mutate_idx = random.randint(1, len(path) - 2)
            planner = UAVPathPlanner(path[mutate_idx - 1], path[-1], self._environment)
            path[mutate_idx] = planner.dodge_obstacle(path[mutate_idx - 1], path[-1])

    def evolve(self):  # GAE:50-58
        self._scores = self._score(self.population)   # one launch instead of P/2 * P evaluations
        new_population = []
        for _ in range(self.population_size // 2):
            parent1, parent2 = self.selection()
            child1, child2 = self.crossover(parent1, parent2)
            self.mutate(child1)
            self.mutate(child2)
            new_population.extend([child1, child2])
        self.population = new_population
        self._scores = None

    def run(self, source, destination, environment):  # GAE:60-66
        self._environment = environment
        self.initialize_population(source, destination, environment)
        best_path = None
        for generation in range(self.generations):
            self.evolve()
            scores = self._score(self.population)
            best = max(range(len(scores)), key=scores.__getitem__)   # max() keeps the first maximum
            best_path = self.population[best]
            print(f"Generation {generation + 1}, Best Fitness: {scores[best]}")
        return best_path


class Environment:  # GAE:87-122 (voxel-set environment; same np.random draw order)
    def __init__(self, x_bounds, y_bounds, z_bounds, num_random_obstacles=30):
        self.x_bounds = x_bounds
        self.y_bounds = y_bounds
        self.z_bounds = z_bounds
        self.num_random_obstacles = num_random_obstacles
        self.obstacles = self.generate_obstacles()

    def generate_obstacles(self):
        obstacles = []
        mid_x = (self.x_bounds[0] + self.x_bounds[1]) // 2
        mid_y = (self.y_bounds[0] + self.y_bounds[1]) // 2
        mid_z = (self.z_bounds[0] + self.z_bounds[1]) // 2
        for dx in range(-3, 4):
            for dy in range(-3, 4):
                for dz in range(-1, 2):
                    obstacles.append((mid_x + dx, mid_y + dy, mid_z + dz))
        for dx in range(-2, 3):
            for dy in range(-2, 3):
                for dz in range(-2, 3):
                    obstacles.append((mid_x + dx + 5, mid_y + dy - 5, mid_z + dz))
        for _ in range(self.num_random_obstacles):
            x = np.random.randint(*self.x_bounds)
            y = np.random.randint(*self.y_bounds)
            z = np.random.randint(*self.z_bounds)
            obstacles.append((x, y, z))
        return list(set(obstacles))


class UAVPathPlanner:  # GAE:125-179 (sequential walk; initial population and the dodge mutation)
    def __init__(self, source, destination, environment):
        self.source = source
        self.destination = destination
        self.environment = environment
        self.path = [self.source]

    def is_within_bounds(self, point):
        x, y, z = point
        return (self.environment.x_bounds[0] <= x <= self.environment.x_bounds[1]
                and self.environment.y_bounds[0] <= y <= self.environment.y_bounds[1]
                and self.environment.z_bounds[0] <= z <= self.environment.z_bounds[1])

    def is_collision(self, point):
        return point in self.environment.obstacles

    def compute_straight_line(self, start, end):
        points = []
        num_steps = max(abs(end[0] - start[0]), abs(end[1] - start[1]), abs(end[2] - start[2]))
        for t in np.linspace(0, 1, num_steps + 1):
            points.append((int(round(start[0] + t * (end[0] - start[0]))), int(round(start[1] + t * (end[1] - start[1]))),
                           int(round(start[2] + t * (end[2] - start[2])))))
        return points

    def dodge_obstacle(self, current, target):
        possible_moves = [(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)]
        random.shuffle(possible_moves)
        for move in possible_moves:
            next_point = (current[0] + move[0], current[1] + move[1], current[2] + move[2])
            if self.is_within_bounds(next_point) and not self.is_collision(next_point):
                return next_point
        return current

    def find_path(self):
        current = self.source
        while current != self.destination:
            straight_line = self.compute_straight_line(current, self.destination)
            for point in straight_line:
                if self.is_collision(point):
                    current = self.dodge_obstacle(current, self.destination)
                    break
                else:
                    self.path.append(point)
                    current = point
                    if current == self.destination:
                        break
        return self.path
