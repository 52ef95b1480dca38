'''
GPU end-to-end tests of the drop-in boundary: ``composer train transformer`` on
a small ``.data`` dataset, resume from its checkpoints, ``evaluate`` and
``generate`` with a ``.data`` and a MIDI prompt, through the click commands.
Also the scaled configuration of BASELINE.json configs[4] (d_model 1024, head
size 64) against the oracle.
'''

import os

import numpy as np
import pytest
from click.testing import CliRunner

pytestmark = pytest.mark.gpu

SMALL_CONFIG = '''
dataset:
    time_step_increment: 10
    max_time_steps: 100
    velocity_bins: 32
transformer:
    model:
        window_size: 64
        embedding_size: 256
        decoder_layers_count: 2
        attention_head_count: 16
        use_relative_attention: false
        attention_dropout_rate: 0.1
        residual_dropout_rate: 0.1
        layer_normalization_epsilon: 0.00001
        scale_attention: true
        initializer_mean: 0
        initializer_stddev: 0.02
        use_layer_normalization: true
    train:
        batch_size: 4
        learning_rate: 0.001
'''


def test_train_resume_evaluate_generate(tmp_path):
    from test_pipeline import _write_dataset
    from composer_b200 import cli as cli_module
    from composer_b200.dataset import sequence

    _write_dataset(tmp_path / 'data' / 'train', files=3, events=900)
    _write_dataset(tmp_path / 'data' / 'test', files=1, events=600, seed=9)
    config_path = tmp_path / 'config.yml'
    config_path.write_text(SMALL_CONFIG)
    logdir = tmp_path / 'logs'
    runner = CliRunner()
    result = runner.invoke(cli_module.cli, ['--seed', '7', 'train', 'transformer', str(tmp_path / 'data'), '--logdir',
                                            str(logdir), '-c', str(config_path), '-e', '3', '--save-freq', '4',
                                            '--no-show-progress-bar'], catch_exceptions=False)
    assert result.exit_code == 0, result.output
    runs = list(logdir.iterdir())
    assert len(runs) == 1 and runs[0].name.startswith('transformer-')
    run = runs[0]
    assert (run / 'config.yml').read_text().startswith('####')       # banner + config backup (cli.py:553-577)
    checkpoints = sorted(p.name for p in run.glob('ckpt-*.npz'))
    assert 1 <= len(checkpoints) <= 3                                  # --max-checkpoints default 3
    # resume: continues from the saved step / epoch counters
    result = runner.invoke(cli_module.cli, ['--seed', '7', 'train', 'transformer', str(tmp_path / 'data'),
                                            '--restoredir', str(run), '-e', '4', '--no-show-progress-bar'],
                           catch_exceptions=False)
    assert result.exit_code == 0, result.output
    result = runner.invoke(cli_module.cli, ['evaluate', 'transformer', str(tmp_path / 'data'), str(run)],
                           catch_exceptions=False)
    assert result.exit_code == 0, result.output
    # generate from a .data prompt and from a MIDI prompt
    prompt = next((tmp_path / 'data' / 'test').glob('*.data'))
    out = tmp_path / 'out' / 'song.mid'
    result = runner.invoke(cli_module.cli, ['--seed', '3', 'generate', 'transformer', str(run), str(out), '-p',
                                            str(prompt), '--prompt-length', '5', '-l', '40'], catch_exceptions=False)
    assert result.exit_code == 0, result.output
    assert out.exists() and out.read_bytes()[:4] == b'MThd'
    midi_prompt = tmp_path / 'prompt.mid'
    sequence.NoteSequence([sequence.Note(0, 400, 60, 90), sequence.Note(400, 900, 67, 70),
                           sequence.Note(900, 1400, 72, 60)]).to_midi(str(midi_prompt))
    out2 = tmp_path / 'out' / 'many.mid'
    result = runner.invoke(cli_module.cli, ['--seed', '3', 'generate', 'transformer', str(run), str(out2), '-p',
                                            str(midi_prompt), '--prompt-length', '6', '-l', '30', '--count', '3'],
                           catch_exceptions=False)
    assert result.exit_code == 0, result.output
    assert sorted(p.name for p in (tmp_path / 'out').glob('many-*.mid')) == ['many-0.mid', 'many-1.mid', 'many-2.mid']
    # the reference's default invocation (--prompt-length 10, --length 1024) asks for more positions than the
    # positional table has rows (window_size 64 here): served in re-primed windows, all 1024 events are written
    out3 = tmp_path / 'out' / 'default_length.mid'
    result = runner.invoke(cli_module.cli, ['--seed', '3', 'generate', 'transformer', str(run), str(out3), '-p',
                                            str(prompt)], catch_exceptions=False)
    assert result.exit_code == 0, result.output
    assert out3.exists() and out3.read_bytes()[:4] == b'MThd'


def test_loss_decreases_when_overfitting_one_batch():
    import kernel_checks
    model, cfg, _ = kernel_checks._small_model(2, 256, 16, window=64, dropout=0.1)
    rng = np.random.default_rng(0)
    draw = rng.integers(0, cfg.vocab_size, size=(4, 65))
    x, y = draw[:, :-1], draw[:, 1:]
    losses = [float(model.train_step(x, y, 1e-3)[0]) / x.size for _ in range(40)]
    assert losses[-1] < 0.5 * losses[0], losses[::8]


def test_scaled_configuration_matches_oracle():
    # BASELINE.json configs[4]: d_model 1024, 16 heads (head size 64); 2 layers keep the oracle quick
    import kernel_checks
    kernel_checks.check_engine_forward_backward(B=1, T=96, layers=2, embedding=1024, heads=16)


def test_head_size_32_matches_oracle():
    import kernel_checks
    kernel_checks.check_engine_forward_backward(B=2, T=80, layers=1, embedding=512, heads=16)


def test_tensorflow_checkpoint_restores_into_the_engine(tmp_path):
    '''
    A log directory in the reference's checkpoint format (tensor bundle + ``checkpoint`` state file, object-graph
    names ``model/decoder_blocks/<i>/...``; here written by ``export_tf_checkpoint``) restores through the same
    ``load_from_checkpoint`` / ``train(restoredir=...)`` entry points as this package's own files: variables,
    Adam slots and counters.
    '''
    import torch
    from composer_b200.models.transformer import Transformer

    def make(seed):
        return Transformer(390, 256, 64, 2, 16, False, 0.0, 0.02, 0.0, 0.0, 1e-5, True, True, seed=seed)

    rng = np.random.default_rng(0)
    draw = rng.integers(0, 390, size=(2, 33))
    source = make(1)
    source.compile(1e-3)
    for _ in range(2):
        source.train_step(draw[:, :-1], draw[:, 1:])
    source._global_step, source._epoch = 3, 2
    source.export_tf_checkpoint(tmp_path / 'logs')
    assert (tmp_path / 'logs' / 'ckpt-1.index').exists() and (tmp_path / 'logs' / 'checkpoint').exists()

    weights_only = make(2)
    weights_only.load_from_checkpoint(tmp_path / 'logs')
    for name, value in source.get_weights().items():
        np.testing.assert_array_equal(weights_only.get_weights()[name], value)
    a, _ = source(draw[:, :-1])
    b, _ = weights_only(draw[:, :-1])
    assert torch.equal(a, b)

    resumed = make(3)
    resumed._restore(Transformer.latest_checkpoint(tmp_path / 'logs'), with_optimizer=True)
    assert (resumed._global_step, resumed._epoch, resumed._adam_t) == (3, 2, 2)
    assert torch.equal(resumed._adam_m, source._adam_m) and torch.equal(resumed._adam_v, source._adam_v)
    source.train_step(draw[:, :-1], draw[:, 1:])
    resumed.compile(1e-3)
    resumed.train_step(draw[:, :-1], draw[:, 1:])
    # (weight gradients are summed with fp32 atomics: the order, hence the last bit, varies from run to run)
    assert float((resumed._params - source._params).abs().max()) < 1e-6
