"""jax.random stand-in: NOT the JAX PRNG (TEST INFRASTRUCTURE ONLY; unused on the hot path)."""


def PRNGKey(seed):
    raise NotImplementedError("the shim does not emulate the JAX PRNG")


split = normal = uniform = randint = permutation = PRNGKey
