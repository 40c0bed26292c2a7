"""Push-style dataflow runtime: the plugin interface of the reference (livenodes/Node.py:22-199), restated.

A graph is wired with `B(...)(A)`; `A.output_data(frame)` then calls `B.add_data(frame)` synchronously and
depth-first on the caller's thread.  Plain callables are accepted as outputs.  `set_passthrough` lets a node
expose an inner sub-graph as itself.  Per-node timing (`activate_timing`) attaches a timestamping Receiver
to every node that gets an output, exactly as the reference does (Node.py:133-140)."""
import collections
import functools

timing_active = False


def activate_timing():
    """Every node that gains an output from now on records [time.time(), frame] for each frame it emits."""
    global timing_active
    timing_active = True


class Node:
    def __init__(self, name="Node", has_inputs=True, has_outputs=True, dont_time=False):
        self.name = name
        self.has_inputs = has_inputs
        self.has_outputs = has_outputs
        self.dont_time = dont_time
        self.input_is_set = False
        self.input_classes = []
        self.output_classes = []
        self.frame_callbacks = []
        self.timing_receiver = None
        self.have_timer = False

    # -- wiring -------------------------------------------------------------------------------
    def __call__(self, input_classes):
        self.set_inputs(input_classes)
        return self

    def set_inputs(self, input_classes):
        if not self.has_inputs:
            raise ValueError("Module does not have inputs.")
        if self.input_is_set:
            raise ValueError("Module input already set.")
        sources = input_classes if isinstance(input_classes, list) else [input_classes]
        for index, source in enumerate(sources):
            source.add_output(self, index)
        self.input_classes = sources
        self.input_is_set = True

    def add_output(self, new_output, data_id=None):
        if timing_active and not self.have_timer and not self.dont_time:
            self.have_timer = True
            from . import Receiver
            self.timing_receiver = Receiver.Receiver(name=self.name + ".Timing", perform_timing=True, dont_time=True)(self)
        if not self.has_outputs:
            raise ValueError("Module does not have outputs.")
        if isinstance(new_output, Node):
            self.output_classes.append(new_output)
            callback = new_output.add_data
        else:
            callback = new_output
        if data_id is not None:
            callback = functools.partial(callback, data_id=data_id)
        self.frame_callbacks.append(callback)

    def set_passthrough(self, node_in, node_out):
        """Make this node a facade for the sub-graph node_in -> ... -> node_out."""
        for attr in ('get_inputs', 'set_inputs', 'add_data', 'start_processing', 'stop_processing'):
            setattr(self, attr, getattr(node_in, attr))
        for attr in ('get_outputs', 'add_output'):
            setattr(self, attr, getattr(node_out, attr))

    def get_inputs(self):
        return self.input_classes

    def get_outputs(self):
        return self.output_classes

    # -- data ---------------------------------------------------------------------------------
    def output_data(self, data_frame):
        for callback in self.frame_callbacks:
            callback(data_frame)

    def add_data(self, data_frame, data_id=0):
        self.output_data(data_frame)

    # -- lifecycle ----------------------------------------------------------------------------
    def start_processing(self, recurse=True):
        if recurse:
            for node in self.output_classes:
                node.start_processing()

    def stop_processing(self, recurse=True):
        if recurse:
            for node in self.output_classes:
                node.stop_processing()

    def get_timing_info(self):
        info = collections.OrderedDict()
        if self.timing_receiver is None:
            return info
        info[self.name] = self.timing_receiver.get_data()
        for node in self.output_classes:
            for child_name, sequence in node.get_timing_info().items():
                info[self.name + "|" + child_name] = sequence
        return info
