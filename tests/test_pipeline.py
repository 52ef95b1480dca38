'''
CPU tests of the rows either side of the hot path: the ``.data`` input pipeline
(reference: composer/models/__init__.py:160-313), MIDI I/O used by ``generate``
(composer/dataset/sequence.py:594-680) and the command-line surface
(composer/cli.py).
'''

import os

import numpy as np
import pytest
from click.testing import CliRunner

from composer_b200 import cli as cli_module
from composer_b200 import config as config_module
from composer_b200 import data
from composer_b200.dataset import midi, sequence

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write_dataset(directory, files=3, events=700, seed=0):
    rng = np.random.default_rng(seed)
    vocabulary = sequence.EventVocabulary(10, 100, 32)
    os.makedirs(directory, exist_ok=True)
    all_ids = []
    for index in range(files):
        ids = rng.integers(0, 390, size=events)
        decoded = [sequence.IntegerEncodedEventSequence.id_to_event(int(i), vocabulary.ranges, vocabulary.value_ranges)
                   for i in ids]
        sequence.EventSequence(decoded, 10, 100, 32).to_integer_encoding().to_file(
            os.path.join(directory, 'piece-%d.data' % index))
        all_ids.append(ids)
    return all_ids


def test_windows_are_shifted_by_one_and_remainders_dropped(tmp_path):
    ids = _write_dataset(tmp_path / 'train', files=2, events=501)
    files = data.get_processed_files(tmp_path / 'train')
    assert len(files) == 2
    events = data.load_events(files)
    assert events.tolist() == np.concatenate(ids).tolist()          # one stream, file order, crossing file borders
    dataset = data.EventWindowDataset(events, batch_size=2, window_size=64, shuffle=False)
    batches = list(dataset)
    assert len(batches) == (1002 // 65) // 2 == len(dataset)
    x, y = batches[0]
    assert x.shape == y.shape == (2, 64) and x.dtype == np.int32
    assert x[0].tolist() == events[:64].tolist() and y[0].tolist() == events[1:65].tolist()
    assert x[1].tolist() == events[65:129].tolist()


def test_shuffle_is_a_permutation_and_changes_each_epoch():
    events = np.arange(33 * 40) % 390
    dataset = data.EventWindowDataset(events, batch_size=4, window_size=32, shuffle=True, seed=5,
                                      shuffle_buffer_batches=2)
    first = np.concatenate([x[:, 0] for x, _ in dataset])
    second = np.concatenate([x[:, 0] for x, _ in dataset])
    assert len(first) == 40 and len(set(first.tolist())) == len(first)
    assert first.tolist() != second.tolist()
    again = data.EventWindowDataset(events, batch_size=4, window_size=32, shuffle=True, seed=5,
                                    shuffle_buffer_batches=2)
    assert np.concatenate([x[:, 0] for x, _ in again]).tolist() == first.tolist()


def test_rank_shards_are_disjoint_and_equal_sized():
    events = np.arange(17 * 23) % 390
    per_rank = []
    for rank in range(3):
        dataset = data.EventWindowDataset(events, batch_size=2, window_size=16, shuffle=True, seed=1, rank=rank,
                                          world_size=3)
        per_rank.append([tuple(x[:, 0].tolist()) for x, _ in dataset])
    assert len({len(batches) for batches in per_rank}) == 1 and len(per_rank[0]) == len(dataset)
    flat = [b for batches in per_rank for b in batches]
    assert len(set(flat)) == len(flat)


def test_empty_and_short_inputs():
    assert list(data.EventWindowDataset(np.zeros(0, dtype=np.uint16), 2, 8)) == []
    assert list(data.EventWindowDataset(np.arange(8), 1, 8)) == []      # needs window + 1 ids
    assert len(list(data.EventWindowDataset(np.arange(9), 1, 8))) == 1


def test_midi_round_trip(tmp_path):
    notes = sequence.NoteSequence(
        [sequence.Note(0, 500, 60, 80), sequence.Note(250, 1000, 64, 100), sequence.Note(1000, 1250, 60, 30)],
        [sequence.SustainPeriod(100, 900)])
    path = tmp_path / 'out.mid'
    notes.to_midi(str(path))
    back = sequence.NoteSequence.from_midi(str(path))
    got = sorted((round(n.start), round(n.end), n.pitch, n.velocity) for n in back.notes)
    assert got == [(0, 500, 60, 80), (250, 1000, 64, 100), (1000, 1250, 60, 30)]
    assert [(round(s.start), round(s.end)) for s in back.sustain_periods] == [(100, 900)]
    # and through the tokenizer
    events = back.to_event_sequence(10, 100, 32)
    assert len(events.events) > 0 and max(events.to_ids()) < 390


def test_midi_reader_handles_tempo_changes_and_running_status(tmp_path):
    # format 1, 96 ppq, tempo 120 bpm then 60 bpm at tick 96; one note from tick 0 to 192 using running status
    tempo_track = b'\x00\xFF\x51\x03\x07\xA1\x20' + b'\x60\xFF\x51\x03\x0F\x42\x40' + b'\x00\xFF\x2F\x00'
    note_track = b'\x00\x90\x3C\x40' + b'\x81\x40\x3C\x00' + b'\x00\xFF\x2F\x00'
    blob = b'MThd' + (6).to_bytes(4, 'big') + (1).to_bytes(2, 'big') + (2).to_bytes(2, 'big') + (96).to_bytes(2, 'big')
    for track in (tempo_track, note_track):
        blob += b'MTrk' + len(track).to_bytes(4, 'big') + track
    path = tmp_path / 'tempo.mid'
    path.write_bytes(blob)
    seq = midi.read_note_sequence(path)
    assert len(seq.notes) == 1
    assert abs(seq.notes[0].start) < 1e-9 and abs(seq.notes[0].end - 1500.0) < 1e-6    # 500 ms + 1000 ms


def test_cli_registry_and_helpers():
    cfg = config_module.get(cli_module.get_default_config())
    assert cli_module._get_event_vocab_size(cfg) == 390
    assert [m.value for m in cli_module.ModelType] == ['music_rnn', 'transformer']
    assert cli_module.get_batch_size(cli_module.ModelType.TRANSFORMER, cfg) == 1
    assert cli_module.get_window_size(cli_module.ModelType.TRANSFORMER, cfg) == 1024
    assert cli_module.get_learning_rate(cli_module.ModelType.TRANSFORMER, cfg) == 0.001
    event = cli_module.decode_to_event(cfg, 388)
    assert event.type == sequence.EventType.SUSTAIN_ON


def test_cli_commands_and_options_match_the_reference(tmp_path):
    runner = CliRunner()
    result = runner.invoke(cli_module.cli, ['--help'])
    assert result.exit_code == 0
    for command in ('train', 'generate', 'evaluate', 'make-config', 'summary', 'preprocess', 'synthesize'):
        assert command in result.output
    help_text = runner.invoke(cli_module.cli, ['train', '--help']).output
    for option in ('--logdir', '--restoredir', '--config', '--epochs', '--use-generator', '--max-files',
                   '--save-freq-mode', '--save-freq', '--max-checkpoints', '--show-progress-bar'):
        assert option in help_text
    help_text = runner.invoke(cli_module.cli, ['generate', '--help']).output
    for option in ('--prompt', '--prompt-length', '--length', '--temperature'):
        assert option in help_text
    target = tmp_path / 'my_config.yml'
    assert runner.invoke(cli_module.cli, ['make-config', str(target)]).exit_code == 0
    assert config_module.get(target).transformer.model.embedding_size == 256
    assert runner.invoke(cli_module.cli, ['-v', 'LOUD', 'make-config', str(target)]).exit_code != 0
    assert runner.invoke(cli_module.cli, ['train', 'lstm', 'x']).exit_code != 0      # unknown model type
    assert runner.invoke(cli_module.cli, ['preprocess', 'a', 'b']).exit_code == 1


def test_generate_requires_a_config_in_restoredir(tmp_path):
    runner = CliRunner()
    result = runner.invoke(cli_module.cli, ['generate', 'transformer', str(tmp_path), str(tmp_path / 'o.mid')])
    assert result.exit_code == 1


def test_get_dataset_errors(tmp_path):
    cfg = config_module.get(cli_module.get_default_config())
    with pytest.raises(cli_module.InvalidParameterError):
        cli_module.get_dataset(cli_module.ModelType.TRANSFORMER, tmp_path, cfg, 'validation')
    with pytest.raises(cli_module.DatasetError):
        cli_module.get_dataset(cli_module.ModelType.TRANSFORMER, tmp_path, cfg, 'train')
