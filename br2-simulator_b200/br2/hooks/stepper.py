"""Stepper seam: ``do_step, stages = extend_stepper_interface(PositionVerlet(), simulator)`` and
``time = do_step(stepper, stages, simulator, time, dt)`` exactly as the reference calls them
(/root/reference/br2/environment.py:155, :221-223, :279-285).

One call advances one PositionVerlet step of every batch instance on the GPU; pass ``n_steps=K`` to
advance K steps in a single kernel launch (state stays in registers between the steps).  Diagnostics
callbacks are the caller's business when K > 1 (``Environment.run`` splits launches at capture steps).
"""


class PositionVerlet:
    """Marker for the only integrator of the path (half kinematic / dynamic / half kinematic)."""


class _DeviceStages:
    def __init__(self, stepper, collection):
        if not isinstance(stepper, PositionVerlet):
            raise TypeError("only PositionVerlet is implemented on the device")
        self.collection = collection


def _do_step(stepper, stages_and_updates, collection, time, dt, n_steps=1, callbacks=True, push_actuation=True):
    engine = collection.engine
    if engine is None:
        raise RuntimeError("finalize() the collection before stepping")
    if push_actuation:
        # the reference's actuation operators read their shared python lists at every step
        # (/root/reference/br2/free_custom_systems.py:49-57, :96-101): upload the current values
        program = getattr(collection, "program", None)
        if program is not None and program.action_lists:
            engine.set_pressures([shared[0] for shared in program.action_lists])
    end = engine.step(n_steps, time, dt)
    if callbacks and collection._callback_ops:
        current_step = round(end / dt)
        if collection.callbacks_due(current_step):
            # collectors that can record from a device-side clone of the batch state do so (no host copy until their
            # histories are read); any other callback object gets the reference's protocol on a host snapshot
            from .._capture import CaptureSink, capture_state

            snap = capture = None
            for rod_index, op in collection._callback_ops:
                lazy = getattr(op, "make_callback_lazy", None)
                if lazy is not None and isinstance(getattr(op, "callback_params", None), CaptureSink):
                    if capture is None:
                        capture = capture_state(engine)
                    lazy(capture, rod_index, collection._systems[rod_index], end, current_step)
                else:
                    if snap is None:
                        snap = collection.snapshot_views()
                    op.make_callback(snap[rod_index], end, current_step)
    return end


def extend_stepper_interface(stepper, system_collection):
    return _do_step, _DeviceStages(stepper, system_collection)
