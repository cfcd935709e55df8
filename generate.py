#!/usr/bin/env python
"""Drop-in for the reference's generate.py (same flags: -restore -audio -speakers -mode -params),
running VQ + WaveNet fast generation on the B200 library instead of TensorFlow.

  python generate.py -restore runs/vctk/weights-110640 -audio p225_001.wav -speakers p225 p226 None -mode sample

Under torchrun (RANK / WORLD_SIZE / LOCAL_RANK set) every rank generates a contiguous slice of the -speakers list on its
own GPU.  Outputs, as the reference writes them (generate.py:94-101,115-117):
  <dir>/embedding_<gs>.npy  <dir>/speaker_embedding_<gs>.npy  <dir>/<gs>_<speaker>.wav (float32, 16 kHz)

Weights: the reference's own TensorFlow checkpoint `<restore>.index` / `<restore>.data-*` (read without TensorFlow,
EMA shadows preferred as `ema.variables_to_restore()` does), or `<restore>.npz` holding arrays keyed by the
reference's variable names.
Encoder: Encoder_64 ("encoder": "64"), Encoder_Magenta ("Magenta") and Encoder_2019 ("2019", hop 320: the trimmed audio
length must also be a multiple of 320) run on the device; -z_e <file.npy> ([F,latent_dim] or [B,F,latent_dim]) supplies
an encoder output instead, -audio then only fixes the length.
"""
import os
import sys
from argparse import ArgumentParser

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main(argv=None):
    import vqvae_wavenet_b200 as pkg
    from vqvae_wavenet_b200 import utils, wavio

    parser = ArgumentParser()
    parser.add_argument('-restore', dest='restore_path', help='path to weights')
    parser.add_argument('-audio', dest='audio_path', help='path to audio')
    parser.add_argument('-speakers', nargs='+', dest='speakers', help='speaker id')
    parser.add_argument('-mode', default='sample', dest='mode', help='decode mode, sample or greedy')
    parser.add_argument('-params', default='model_parameters.json', dest='parameter_path', metavar='str',
                        help='path to parameters file')
    parser.add_argument('-z_e', dest='z_e_path', default=None, help='encoder output (.npy)')
    parser.add_argument('-device', dest='device', type=int, default=0)
    parser.add_argument('-seed', dest='seed', type=int, default=None, help='seed of the draw stream (sample mode)')
    parser.add_argument('-precision', dest='precision', default='fp32', choices=('fp32', 'tc', 'bf16'),
                        help='fp32: float32 contractions (the reference arithmetic); bf16: tcgen05 tensor-core kernel')
    args = parser.parse_args(argv)

    gs = int(args.restore_path.split('-')[-1])                     # generate.py:33
    save_path = args.restore_path.split('/weights')[0]
    dataset, num_speakers = utils.dataset_for_speakers(args.speakers)   # generate.py:46-57 (decided on the full list)

    # one process per GPU (torchrun / any launcher that sets RANK, WORLD_SIZE, LOCAL_RANK): every rank takes a contiguous
    # slice of the requested speakers and writes that slice's WAVs - streams never interact, nothing is exchanged
    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    lo = 0
    if world > 1:
        from vqvae_wavenet_b200 import sharding
        lo, hi = sharding.stream_slice(len(args.speakers), rank, world)
        args.speakers = args.speakers[lo:hi]
        if args.device == 0:
            args.device = int(os.environ.get('LOCAL_RANK', '0'))
        if not args.speakers:
            return
    batch_size = len(args.speakers)

    table = utils.get_speaker_to_int(utils.find_speaker_table(dataset, roots=(".", ROOT)))
    speaker = utils.speaker_onehot(args.speakers, table, num_speakers)  # [B,1,N]

    params_path = args.parameter_path if os.path.exists(args.parameter_path) else os.path.join(ROOT, args.parameter_path)
    cfg = pkg.EngineConfig.from_files(params_path, num_speakers=num_speakers)
    if cfg.model['encoder'] not in ('Magenta', '64', '2019'):
        raise NotImplementedError("encoder %s not implemented" % cfg.model['encoder'])   # generate.py:69 (Q15)

    wav = wavio.prepare_audio(wavio.read_wav(args.audio_path, 16000), batch_size)        # generate.py:36-44
    length = wav.shape[1]
    z_e = None
    if cfg.model['encoder'] == '2019' and args.z_e_path is None and length % 320 != 0:
        # the reference fails inside add_condition's reshape for such lengths; say why
        raise ValueError("Encoder_2019 needs an audio length that is a multiple of 320 samples after the 512-trim, got %d" % length)
    if args.z_e_path is not None:
        z_e = np.load(args.z_e_path).astype(np.float32)
        if z_e.ndim == 2:
            z_e = np.tile(z_e[None], (batch_size, 1, 1))
        if length % z_e.shape[1] != 0:
            raise ValueError("audio length %d is not a multiple of the %d encoder frames" % (length, z_e.shape[1]))

    from vqvae_wavenet_b200 import tf_checkpoint
    weights_file = args.restore_path + '.npz'
    if not tf_checkpoint.is_bundle(args.restore_path) and not os.path.exists(weights_file):
        raise FileNotFoundError("neither a TensorFlow checkpoint (%s.index) nor %s (arrays keyed by reference "
                                "variable name) exists" % (args.restore_path, weights_file))
    engine = pkg.Engine(cfg, device=args.device, max_batch=batch_size)
    engine.set_precision(args.precision)
    wanted = dict((n, s) for n, s, _ in engine.tensor_table())
    if tf_checkpoint.is_bundle(args.restore_path):
        # generate.py:88-90: Saver(ema.variables_to_restore()).restore(sess, restore_path), read without TensorFlow
        for key, value in tf_checkpoint.generator_weights(args.restore_path, wanted).items():
            engine.set_tensor(key, value)
    else:
        with np.load(weights_file) as data:
            for name in data.files:
                key = name[:-len('/ExponentialMovingAverage')] if name.endswith('/ExponentialMovingAverage') else name
                key = key[len('optimiser/'):] if key.startswith('optimiser/') else key      # EMA shadow scope
                if key in wanted:
                    engine.set_tensor(key, data[name])

    encoder = None
    if z_e is None:
        # generate.py:40 tiles ONE utterance over the batch: encode it once, tile the result
        enc_cls = {'64': pkg.Encoder_64, 'Magenta': pkg.Encoder_Magenta, '2019': pkg.Encoder_2019}[cfg.model['encoder']]   # generate.py:65-69
        encoder = enc_cls(cfg.model['latent_dim'], engine)
        z_e = np.tile(encoder.build(wav[:1]), (batch_size, 1, 1))
    model = pkg.VQVAE({'x': wav, 'z_e': z_e, 'speaker': speaker, 'encoder': encoder,
                       'decoder': pkg.WavenetDecoder(cfg.wavenet), 'k': cfg.model['k'], 'beta': cfg.model['beta'],
                       'verbose': cfg.model.get('verbose', False), 'use_vq': cfg.model['use_vq'],
                       'speaker_embedding': cfg.model['speaker_embedding'], 'num_speakers': num_speakers,
                       'engine': engine})
    model.build_generator()
    wavenet = model.decoder.wavenet
    encoding = model.encoding                                                       # generate.py:92

    if cfg.model['use_vq'] and rank == 0:
        np.save(save_path + '/embedding_%d.npy' % gs, model.embedding)              # generate.py:96-98
    if cfg.model['speaker_embedding'] > 0 and rank == 0:
        np.save(save_path + '/speaker_embedding_%d.npy' % gs, model.speaker_embedding)

    uniforms = None
    if args.mode == 'sample' and args.seed is None:
        uniforms = np.random.rand(length, batch_size)                               # utils.py:22, one draw per step
    engine.set_stream_offset(lo)          # -seed draws are keyed on the position in the full -speakers list
    to_write, _ = wavenet.generate(encoding, length, mode=args.mode, uniforms=uniforms,
                                   seed=0 if args.seed is None else args.seed)      # generate.py:103-113
    for i, s in enumerate(args.speakers):
        s = 'no_speaker' if s == 'None' else s
        wavio.write_wav_float32(save_path + '/%d_%s.wav' % (gs, s), 16000, to_write[i])  # generate.py:115-117
    engine.close()


if __name__ == '__main__':
    main()
