from . import brachi, car, quad  # noqa: F401

REGISTRY = {"car": car.define, "brachi": brachi.define, "quad": quad.define}
