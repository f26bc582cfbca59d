from . import car, brachi  # noqa: F401

REGISTRY = {"car": car.define, "brachi": brachi.define}
