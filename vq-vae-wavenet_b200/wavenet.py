"""Mirror of Decoder/WaveNet/wavenet.py's generator-facing surface (wavenet.py:10-21,103-172).

In the reference `build_generator` builds a TF graph and leaves handles (`init_ops`, `push_ops`,
`predictions`, `input_t`, `local_condition_t`) that generate.py evaluates with sess.run.  Here the
same names are callables bound to one vqwn Engine (one GPU)."""
import json


class Wavenet:
    def __init__(self, args_file="wavenet_parameters.json"):
        if isinstance(args_file, dict):
            args = dict(args_file)
        else:
            with open(args_file) as f:
                args = json.load(f)
        assert len(args["dilation_rates"]) == args["num_cycles"] * args["num_cycle_layers"]
        kernel_size = args["kernel_size"]
        self.receptive_field = sum(args["dilation_rates"]) * (kernel_size - 1) + 1
        self.receptive_field += args["preprocess"]["kernel_size"] - 1
        self.args = args
        self._print = (lambda s, t: print(s, t)) if args.get("verbose") else (lambda s, t: None)
        self._print("wavenet receptive_field:", self.receptive_field)
        self.engine = None
        self.batch_size = None

    def build_generator(self, engine, batch_size):
        """binds the one-step generator to a device handle (wavenet.py:103-172)."""
        self.engine = engine
        self.batch_size = batch_size

    # sess.run(wavenet.init_ops)  (generate.py:105)
    def init_ops(self):
        self.engine.reset(self.batch_size)

    # sess.run([wavenet.predictions, wavenet.push_ops], {input_t: audio, local_condition_t: lc})
    def predictions(self, input_t, local_condition_t, return_logits=False):
        probs, logits = self.engine.step(input_t, local_condition_t)
        return (probs, logits) if return_logits else probs

    # init_ops + the whole sample loop of generate.py:103-113 in one persistent kernel
    def generate(self, encoding, length, mode="sample", uniforms=None, seed=0):
        return self.engine.generate(encoding, length, mode=mode, uniforms=uniforms, seed=seed)
