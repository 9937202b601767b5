"""``elastica.callback_functions`` (TEST INFRASTRUCTURE).  Base of
``FreeCallback`` (/root/reference/br2/free_simulator.py:42-73)."""


class CallBackBaseClass:
    def __init__(self):
        pass

    def make_callback(self, system, time, current_step):
        pass
