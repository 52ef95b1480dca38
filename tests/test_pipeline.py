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


# ---------------------------------------------------------------------------
# `.data` codec against files written by the reference (sequence.py:1500-1587)
# ---------------------------------------------------------------------------

def _data_golden():
    import json
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'data_golden.json')
    with open(path) as handle:
        fixture = json.load(handle)
    import sys
    assert fixture['byteorder'] == sys.byteorder      # the reference packs with native byte order
    return fixture['cases']


def test_data_files_match_the_reference_writer_byte_for_byte(tmp_path):
    '''Our writer reproduces the reference's `.data` bytes; our readers return what the reference's readers do.'''

    cases = _data_golden()
    assert len(cases) >= 10 and any(not case['ids'] for case in cases)        # includes an empty sequence
    for index, case in enumerate(cases):
        reference_bytes = bytes.fromhex(case['file_hex'])
        notes = [sequence.Note(*n) for n in case['notes']]
        sustains = [sequence.SustainPeriod(*s) for s in case['sustain_periods']]
        events = sequence.NoteSequence(notes, sustains).to_event_sequence(
            case['time_step_increment'], case['max_time_steps'], case['velocity_bins'])
        ours = tmp_path / ('ours_%d.data' % index)
        events.to_integer_encoding().to_file(str(ours))
        assert ours.read_bytes() == reference_bytes, 'case %d: written bytes differ from the reference writer' % index
        theirs = tmp_path / ('reference_%d.data' % index)
        theirs.write_bytes(reference_bytes)
        ids, _, ranges, settings = sequence.IntegerEncodedEventSequence.event_ids_from_file(str(theirs))
        assert list(ids) == case['ids']
        assert tuple(settings) == (case['time_step_increment'], case['max_time_steps'], case['velocity_bins'])
        as_array = sequence.IntegerEncodedEventSequence.event_ids_from_file(str(theirs), as_numpy_array=True)[0]
        assert as_array.tolist() == case['ids']
        assert list(sequence.IntegerEncodedEventSequence.event_ids_from_file_as_generator(str(theirs))) == case['ids']
        back = sequence.IntegerEncodedEventSequence.from_file(str(theirs), decode=False)
        assert [list(pair) for pair in back.events] == case['pairs']
        # decode=True gives the event sequence the file was written from
        decoded = sequence.IntegerEncodedEventSequence.from_file(str(theirs), decode=True)
        assert [(e.type, e.value) for e in decoded.events] == [(e.type, e.value) for e in events.events]


def test_data_files_against_the_live_reference_codec(tmp_path):
    '''Both directions against the reference module itself when /root/reference is present (skipped elsewhere).'''

    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
    from reference_shims import load_reference_sequence
    ref = load_reference_sequence()
    if ref is None:
        pytest.skip('the reference tree is not available on this machine')
    rng = np.random.default_rng(5)
    vocabulary = sequence.EventVocabulary(10, 100, 32)
    for index in range(6):
        ids = rng.integers(0, 390, size=int(rng.integers(0, 300)))
        decoded = [sequence.IntegerEncodedEventSequence.id_to_event(int(i), vocabulary.ranges, vocabulary.value_ranges)
                   for i in ids]
        ours = tmp_path / ('ours_%d.data' % index)
        sequence.EventSequence(decoded, 10, 100, 32).to_integer_encoding().to_file(str(ours))
        # the reference reads our file ...
        ref_ids = [int(v) for v in ref.IntegerEncodedEventSequence.event_ids_from_file(str(ours))[0]]
        assert ref_ids == ids.tolist()
        # ... and re-writes it identically
        theirs = tmp_path / ('theirs_%d.data' % index)
        ref.IntegerEncodedEventSequence.from_file(str(ours), decode=False).to_file(str(theirs))
        assert theirs.read_bytes() == ours.read_bytes()
