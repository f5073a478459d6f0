"""sample_factory.launcher.launcher_utils: seeds(n) used by experiment definition files"""
import random


def seeds(num_seeds: int):
    return [random.randrange(1000000, 9999999) for _ in range(num_seeds)]
