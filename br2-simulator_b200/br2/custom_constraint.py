"""``LastEndFixedRod`` (reference: /root/reference/br2/custom_constraint.py:9-88): hard-pins the LAST
node's position and the LAST element's director, and zeroes their velocity / omega."""
from .hooks.ops import CONS_LAST_END_FIXED, ConstraintBase


class LastEndFixedRod(ConstraintBase):
    def __init__(self, fixed_position, fixed_directors, **kwargs):
        super().__init__(**kwargs)
        self.fixed_position = fixed_position
        self.fixed_directors = fixed_directors

    def lower_constraint(self, ctx, rod_index):
        ctx.add_constraint(rod_index, CONS_LAST_END_FIXED, self.fixed_position, self.fixed_directors)
