"""Mirror of model.py's VQVAE as generate.py uses it (model.py:7-33,133-159).

`args` keeps the reference's keys (generate.py:71-82): 'x' [B,T,1] audio (or 'z_e' [B,F,D] when the
encoder output is supplied directly), 'speaker' [B,1,N] one-hot, 'encoder', 'decoder', 'k', 'beta',
'verbose', 'use_vq', 'speaker_embedding', 'num_speakers'; plus 'engine' (the device handle that
replaces tf.Session + variables)."""
import numpy as np


class VQVAE:
    def __init__(self, args):
        self.x = args.get("x")
        self.h = args.get("speaker")
        self.encoder = args.get("encoder")
        self.decoder = args["decoder"]
        self.k = args["k"]
        self.beta = args.get("beta", 0.25)
        self.use_vq = args["use_vq"]
        self.num_speakers = args["num_speakers"]
        self.engine = args["engine"]
        self._z_e = args.get("z_e")
        self._print = (lambda s, t: print(s, np.shape(t))) if args.get("verbose") else (lambda s, t: None)
        if self.x is not None:
            self._print("input x:", self.x)
        self.speaker_idx = None
        if self.h is not None:
            # model.py:22: tf.argmax over the one-hot; the all-zero 'None' row gives index 0
            self.speaker_idx = np.argmax(np.asarray(self.h), axis=-1).reshape(-1).astype(np.int32)
            self._print("input h:", self.h)
        self.encoding = None
        self.q_z_x = None

    def build_generator(self):
        """model.py:154-159: encoder -> VQ -> decoder.build_generator."""
        if self._z_e is None:
            if self.encoder is None:
                raise NotImplementedError("no encoder given and no z_e supplied")
            self._z_e = self.encoder.build(self.x)
        self.z_e = np.ascontiguousarray(self._z_e, dtype=np.float32)
        self._print("z_e:", self.z_e)
        spk = self.speaker_idx
        if spk is None:
            spk = np.zeros(self.z_e.shape[0], dtype=np.int32)
        self.q_z_x, self.encoding = self.decoder.build_generator(self.engine, self.z_e, spk)
        if self.q_z_x is not None:
            self._print("q(z|x):", self.q_z_x)

    @property
    def embedding(self):                      # sess.run(model.embedding), generate.py:96-98
        return self.engine.get_tensor("embedding/embedding")

    @property
    def speaker_embedding(self):              # generate.py:99-101
        return self.engine.get_tensor("speaker_embedding")
