"""CartPole (Barto, Sutton & Anderson 1983) with the constants and termination rule of the classic-control task:
gravity 9.8, cart 1.0 kg, pole 0.1 kg / half-length 0.5 m, force 10 N, Euler steps of 0.02 s; the episode terminates when
|x| > 2.4 or |theta| > 12 degrees; reward 1 per step; the start state is uniform in [-0.05, 0.05]^4."""
from __future__ import annotations

import math

import numpy as np

from .. import spaces
from ..core import Env


class CartPoleEnv(Env):
    metadata = {"render_modes": [], "render_fps": 50}

    def __init__(self, render_mode=None):
        self.gravity, self.masscart, self.masspole = 9.8, 1.0, 0.1
        self.total_mass = self.masspole + self.masscart
        self.length = 0.5
        self.polemass_length = self.masspole * self.length
        self.force_mag, self.tau = 10.0, 0.02
        self.theta_threshold_radians = 12 * 2 * math.pi / 360
        self.x_threshold = 2.4
        high = np.array([self.x_threshold * 2, np.finfo(np.float32).max, self.theta_threshold_radians * 2,
                         np.finfo(np.float32).max], dtype=np.float32)
        self.action_space = spaces.Discrete(2)
        self.observation_space = spaces.Box(-high, high, dtype=np.float32)
        self.render_mode = render_mode
        self.state = None

    def reset(self, *, seed=None, options=None):
        super().reset(seed=seed)
        self.state = self.np_random.uniform(low=-0.05, high=0.05, size=(4,))
        return np.array(self.state, dtype=np.float32), {}

    def step(self, action):
        assert self.state is not None, "Call reset before using step method."
        x, x_dot, theta, theta_dot = self.state
        force = self.force_mag if int(action) == 1 else -self.force_mag
        costheta, sintheta = math.cos(theta), math.sin(theta)
        temp = (force + self.polemass_length * theta_dot ** 2 * sintheta) / self.total_mass
        thetaacc = (self.gravity * sintheta - costheta * temp) / (
            self.length * (4.0 / 3.0 - self.masspole * costheta ** 2 / self.total_mass))
        xacc = temp - self.polemass_length * thetaacc * costheta / self.total_mass
        x, x_dot = x + self.tau * x_dot, x_dot + self.tau * xacc
        theta, theta_dot = theta + self.tau * theta_dot, theta_dot + self.tau * thetaacc
        self.state = (x, x_dot, theta, theta_dot)
        terminated = bool(x < -self.x_threshold or x > self.x_threshold or theta < -self.theta_threshold_radians
                          or theta > self.theta_threshold_radians)
        return np.array(self.state, dtype=np.float32), 1.0, terminated, False, {}
