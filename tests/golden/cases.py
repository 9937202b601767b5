"""Shared description of the golden trajectory cases (used by make_golden.py and the tests)."""
PSI = 6895.0

# (name, assembly, reset kwargs, action, duration, rendering_fps)  -- dt stays the default 2e-5
TRAJECTORIES = (
    ("single_gravity", "single", {"gravity": True}, {"action1": 30 * PSI, "action2": 30 * PSI}, 0.08, 25),
    ("single_twist", "single", {}, {"action1": 5 * PSI, "action2": 35 * PSI, "action3": 12 * PSI}, 0.03, 100),
    ("double", "double", {}, {"action1": 25 * PSI, "action2": 10 * PSI, "action3": 20 * PSI}, 0.02, 100),
    ("triple", "triple", {"gravity": True, "nu_multiplier": 2.0, "t_multiplier": 0.5},
     {"action1": 15 * PSI, "action2": 30 * PSI, "action3": 5 * PSI}, 0.01, 200),
)
STATE_FIELDS = ("position_collection", "velocity_collection", "director_collection", "omega_collection",
                "acceleration_collection", "alpha_collection")
CALLBACK_KEYS = ("time", "step", "position", "velocity", "acceleration", "omega", "director",
                 "external_forces", "external_torques", "internal_forces", "internal_torques",
                 "lengths", "dilatation", "radius", "com")
