"""sample_factory.algo.utils.gymnasium_utils (algo/utils/gymnasium_utils.py:96): old-gym -> gymnasium space conversion."""
import gymnasium


def convert_space(space):
    """Spaces of the legacy `gym` package re-created as gymnasium spaces (gymnasium spaces pass through)."""
    sp = gymnasium.spaces
    if isinstance(space, sp.Space):
        return space
    name = type(space).__name__
    if name == "Box":
        return sp.Box(low=space.low, high=space.high, shape=space.shape, dtype=space.dtype)
    if name == "Discrete":
        return sp.Discrete(n=space.n)
    if name == "MultiDiscrete":
        return sp.MultiDiscrete(nvec=space.nvec)
    if name == "Tuple":
        return sp.Tuple(spaces=tuple(convert_space(s) for s in space.spaces))
    if name == "Dict":
        return sp.Dict(spaces={k: convert_space(v) for k, v in space.spaces.items()})
    raise NotImplementedError(f"Cannot convert space of type {space}. Please upgrade your code to gymnasium.")


def patch_non_gymnasium_env(env):
    env.observation_space = convert_space(env.observation_space)
    env.action_space = convert_space(env.action_space)
    return env
