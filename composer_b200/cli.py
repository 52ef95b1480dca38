'''
The command-line interface of the B200 build of Composer.

Same group, commands, arguments and options as the reference for the
Transformer path (composer/cli.py:41-59 group, :516-589 ``train``, :591-615
``evaluate``, :617-680 ``generate``, :69-78 ``make-config``, :424-440
``summary``), the same ``ModelType`` registry and ``create_model`` factory
(:80-141), and the same helper getters (:143-183, :382-422).  Commands of the
reference that are outside the hot path (``preprocess``, ``export-dataset``,
``visualize-training``, ``synthesize``) are kept as entries that explain they
are not provided.

Multi-GPU: launch with ``torchrun --nproc-per-node N -m composer_b200 train ...``;
every rank trains on its shard of the batches and gradients are all-reduced
with NCCL.  ``generate --count K`` shards K independent sequences over ranks.
'''

import datetime
import logging
import os
import time
from enum import Enum, unique
from pathlib import Path
from shutil import copy2

import click
import numpy as np

import composer_b200.config
import composer_b200.logging_utils as logging_utils
from composer_b200 import ModelSaveFrequencyMode
from composer_b200.click_utils import EnumType
from composer_b200.dataset.sequence import (EventSequence, IntegerEncodedEventSequence, NoteSequence,
                                            OneHotEncodedEventSequence)
from composer_b200.exceptions import DatasetError, InvalidParameterError


def _set_verbosity_level(logger, value):
    level = getattr(logging, value.upper(), None)
    if level is None:
        raise click.BadParameter('Must be CRITICAL, ERROR, WARNING, INFO, or DEBUG, not \'{}\''.format(value))

    logger.setLevel(level)


@click.group()
@click.option('--verbosity', '-v', default='INFO', help='Either CRITICAL, ERROR, WARNING, INFO, or DEBUG.')
@click.option('--seed', type=int, help='Sets the seed of the random engine.')
@click.pass_context
def cli(ctx, verbosity, seed):
    '''
    A deep learning enabled music generator (B200 build: Transformer path).

    '''

    if seed is None:
        # Same time-derived default as the reference (cli.py:51-57).  The reference computes the seed
        # and never applies it; here it keys the Philox streams (dropout, sampling) and numpy.
        t = int(time.time() * 1000.0)
        seed = ((t & 0xff000000) >> 24) + ((t & 0x00ff0000) >> 8) + ((t & 0x0000ff00) << 8) + ((t & 0x000000ff) << 24)

    ctx.ensure_object(dict)
    ctx.obj['seed'] = int(seed)
    np.random.seed(seed % (2 ** 32))
    logging_utils.init()
    _set_verbosity_level(logging.getLogger(), verbosity)


def get_default_config():
    '''The default configuration file shipped with the package.'''

    return Path(__file__).parent / 'default_config.yml'


@cli.command()
@click.argument('filepath')
def make_config(filepath):
    '''
    Creates a configuration file from the default configuration.

    '''

    copy2(get_default_config(), filepath)


@unique
class ModelType(Enum):
    '''The type of the model (same members as the reference, cli.py:80-93).'''

    MUSIC_RNN = 'music_rnn'
    TRANSFORMER = 'transformer'


def get_event_sequence_ranges(config):
    '''Event value ranges, dimensions and id ranges for the dataset settings of ``config`` (cli.py:382-398).'''

    event_value_ranges = EventSequence._compute_event_value_ranges(
        config.dataset.time_step_increment, config.dataset.max_time_steps, config.dataset.velocity_bins)
    event_dimensions = EventSequence._compute_event_dimensions(event_value_ranges)
    event_ranges = EventSequence._compute_event_ranges(event_dimensions)
    return event_value_ranges, event_dimensions, event_ranges


def _get_event_vocab_size(config):
    _, _, event_ranges = get_event_sequence_ranges(config)
    return OneHotEncodedEventSequence.get_one_hot_size(event_ranges)


def decode_to_event(config, event_id):
    event_value_ranges, _, event_ranges = get_event_sequence_ranges(config)
    return IntegerEncodedEventSequence.id_to_event(event_id, event_ranges, event_value_ranges)


def create_model(model_type, config, **kwargs):
    '''
    Creates the model registered for ``model_type`` from ``config`` and returns
    ``(model, event_vocab_size)`` (cli.py:95-141).  ``kwargs`` (``device``,
    ``seed``, ``process_group``) go to the model constructor.
    '''

    from composer_b200 import models
    event_vocab_size = _get_event_vocab_size(config)

    def _create_music_rnn():
        return models.MusicRNN()

    def _create_transformer():
        return models.Transformer(
            event_vocab_size, config.transformer.model.embedding_size,
            config.transformer.model.window_size, config.transformer.model.decoder_layers_count,
            config.transformer.model.attention_head_count, config.transformer.model.use_relative_attention,
            config.transformer.model.initializer_mean, config.transformer.model.initializer_stddev,
            config.transformer.model.attention_dropout_rate, config.transformer.model.residual_dropout_rate,
            config.transformer.model.layer_normalization_epsilon, config.transformer.model.scale_attention,
            config.transformer.model.use_layer_normalization, **kwargs
        )

    function_map = {
        ModelType.MUSIC_RNN: _create_music_rnn,
        ModelType.TRANSFORMER: _create_transformer
    }

    return function_map[model_type](), event_vocab_size


def get_batch_size(model_type, config):
    if model_type == ModelType.MUSIC_RNN:
        return config.music_rnn.train.batch_size
    elif model_type == ModelType.TRANSFORMER:
        return config.transformer.train.batch_size
    else:
        raise NotImplementedError('Unrecognized model type: \'{}\'.'.format(model_type))


def get_learning_rate(model_type, config):
    if model_type == ModelType.MUSIC_RNN:
        return config.music_rnn.train.learning_rate
    elif model_type == ModelType.TRANSFORMER:
        return config.transformer.train.learning_rate
    else:
        raise NotImplementedError('Unrecognized model type: \'{}\'.'.format(model_type))


def get_window_size(model_type, config):
    if model_type == ModelType.MUSIC_RNN:
        return config.music_rnn.model.window_size
    elif model_type == ModelType.TRANSFORMER:
        return config.transformer.model.window_size
    else:
        raise NotImplementedError('Unrecognized model type: \'{}\'.'.format(model_type))


def _distributed_context():
    '''(rank, world_size, local_rank); initialises NCCL when launched by torchrun.'''

    world_size = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    import torch
    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    if world_size > 1 and not torch.distributed.is_initialized():
        torch.distributed.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    return rank, world_size, local_rank


def _agreed_seed(ctx):
    '''The run's seed, identical on every rank (rank 0's: without --seed each process draws its own from the clock).'''

    from composer_b200 import parallel
    seed = parallel.agree_on_seed((ctx.obj or {}).get('seed', 0))
    if ctx.obj is not None:
        ctx.obj['seed'] = seed
    np.random.seed(seed % (2 ** 32))
    return seed


def get_dataset(model_type, dataset_path, config, mode='', use_generator=True, max_files=None,
                show_progress_bar=True, shuffle_files=True, shuffle_dataset=True, seed=0, rank=0, world_size=1):
    '''
    Loads the ``mode`` split (``train`` / ``test``) of a preprocessed dataset
    directory as an iterable of ``(x, y)`` batches (cli.py:185-276).  TFRecord
    exports of the reference need TensorFlow and are refused.
    '''

    from composer_b200 import data

    if mode not in ['train', 'test', '']:
        raise InvalidParameterError(
            '\'{}\' is an invalid dataset mode! Must be one of: \'train\', \'test\', or none.'.format(mode))

    dataset_path = Path(dataset_path)
    if dataset_path.is_dir():
        dataset_path = dataset_path / mode
        if not dataset_path.exists():
            raise DatasetError('Could not get {mode} dataset since the specified dataset directory, \'{path}\', '
                               'has no {mode} folder.'.format(path=dataset_path, mode=mode))

        files = data.get_processed_files(dataset_path)
        if shuffle_files:
            np.random.default_rng(seed).shuffle(files)      # same order on every rank
    else:
        raise InvalidParameterError(
            '\'{}\' is an invalid dataset path! The dataset must be a directory of processed MIDI files '
            '(TFRecord exports require TensorFlow and are not supported by the B200 build).'.format(dataset_path))

    if max_files is not None:
        files = files[:max_files]

    dataset = data.load_dataset(files, get_batch_size(model_type, config), get_window_size(model_type, config),
                                show_loading_progress_bar=show_progress_bar, shuffle=shuffle_dataset, seed=seed,
                                rank=rank, world_size=world_size)
    return data.prefetch_pinned(dataset)


def get_config_from_restoredir(restoredir):
    '''The configuration saved next to a model's checkpoints (cli.py:500-514).'''

    config_filepath = Path(restoredir) / 'config.yml'
    if not config_filepath.exists():
        logging.error('Failed to restore model from \'{}\'! Could not find \'config.yml\' file!'.format(restoredir))
        exit(1)

    return composer_b200.config.get(config_filepath)


_CONFIG_COPY_FORMAT = '\n'.join(line.strip() for line in '''
#########################################################
# Datetime: {datetime}.
#########################################################
# This is an autogenerated backup of the configuration file
# used when invoking the train command.
#
# DO NOT MODIFY THIS FILE!
# Doing so may cause errors upon resuming training.
#########################################################
{config_source}
'''.strip().split('\n'))


@cli.command()
@click.argument('model-type', type=EnumType(ModelType, False))
@click.argument('dataset-path')
@click.option('--logdir', default='./output/logdir/', help='The root log directory. Defaults to \'./output/logdir\'.')
@click.option('--restoredir', default=None, type=str, help='The directory of the model to continue training.')
@click.option('-c', '--config', 'config_filepath', default=None,
              help='The path to the model configuration file. If unspecified, uses the default config for the model.' +
              '\n\nIf a restoredir is specified, the configuration file in the restoredir is used instead (and this value is ignored).')
@click.option('-e', '--epochs', 'epochs', default=10, help='The number of epochs to train for. Defaults to 10.')
@click.option('--use-generator/--no-use-generator', default=False,
              help='Accepted for compatibility; the dataset is always loaded into host memory. Defaults to False.')
@click.option('--max-files', default=None, help='The maximum number of files to load. Defaults to None, which means ' +
              'that ALL files will be loaded.', type=int)
@click.option('--save-freq-mode', 'save_frequency_mode', type=EnumType(ModelSaveFrequencyMode, False),
              help='The units of the save frequency. Defaults to GLOBAL_STEP.', default='global_step')
@click.option('--save-freq', 'save_frequency', help='The frequency at which to save the model (in the units specified ' +
              'by the save frequency mode). Defaults to every 500 global steps.', type=int, default=500)
@click.option('--max-checkpoints', 'max_checkpoints', help='The maximum number of checkpoints to keep. Defaults to 3.',
              type=int, default=3)
@click.option('--show-progress-bar/--no-show-progress-bar', 'show_progress_bar', help='Indicates whether a progress bar ' +
              'will be shown to indicate epoch status. Defaults to True.', default=True)
@click.option('--max-steps', default=None, type=int, help='Stop after this many steps (extension, for smoke runs).')
@click.pass_context
def train(ctx, model_type, dataset_path, logdir, restoredir, config_filepath, epochs,
          use_generator, max_files, save_frequency_mode, save_frequency,
          max_checkpoints, show_progress_bar, max_steps):
    '''
    Trains the specified model.

    '''

    rank, world_size, local_rank = _distributed_context()
    seed = _agreed_seed(ctx)
    if restoredir is not None:
        config = get_config_from_restoredir(restoredir)
        model_logdir = None
    else:
        stamp = datetime.datetime.now().strftime('%Y-%m-%d_%H-%M-%S')
        if world_size > 1:
            import torch
            # every rank must agree on the directory name
            holder = [stamp]
            torch.distributed.broadcast_object_list(holder, src=0)
            stamp = holder[0]
        model_logdir = Path(logdir) / '{}-{}'.format(model_type.name.lower(), stamp)
        config = composer_b200.config.get(config_filepath or get_default_config())
        if rank == 0:
            model_logdir.mkdir(parents=True, exist_ok=True)
            with open(config.filepath) as original_config_file, \
                    open(model_logdir / 'config.yml', 'w+') as copy_config_file:
                copy_config_file.write(_CONFIG_COPY_FORMAT.format(datetime=str(datetime.datetime.now()),
                                                                  config_source=original_config_file.read()))

    model, _ = create_model(model_type, config, seed=seed)

    input_shape = (get_batch_size(model_type, config), get_window_size(model_type, config))
    learning_rate = get_learning_rate(model_type, config)
    train_dataset = get_dataset(model_type, dataset_path, config, 'train', use_generator, max_files=max_files,
                                show_progress_bar=show_progress_bar and rank == 0, seed=seed, rank=rank,
                                world_size=world_size)
    model.train(
        train_dataset, input_shape, model_logdir, restoredir=restoredir, epochs=epochs,
        learning_rate=learning_rate, save_frequency_mode=save_frequency_mode,
        save_frequency=save_frequency, max_checkpoints=max_checkpoints,
        show_progress_bar=show_progress_bar, max_steps=max_steps
    )


@cli.command()
@click.argument('model-type', type=EnumType(ModelType, False))
@click.argument('dataset-path')
@click.argument('restoredir')
@click.option('--use-generator/--no-use-generator', default=False, help='Accepted for compatibility.')
@click.option('--max-files', default=None, help='The maximum number of files to load. Defaults to None, which means ' +
              'that ALL files will be loaded.', type=int)
def evaluate(model_type, dataset_path, restoredir, use_generator, max_files):
    '''
    Evaluate the specified model.

    '''

    config = get_config_from_restoredir(restoredir)
    model, _ = create_model(model_type, config)
    model.load_from_checkpoint(restoredir)

    model.compile(get_learning_rate(model_type, config))
    model.build(input_shape=(get_batch_size(model_type, config), None))

    test_dataset = get_dataset(model_type, dataset_path, config, 'test', use_generator, max_files=max_files,
                               shuffle_dataset=False)
    loss, accuracy = model.evaluate(test_dataset, verbose=0)
    logging.info('- Finished evaluating model. Loss: {:.4f}, Accuracy: {:.4f}'.format(loss, accuracy))


@cli.command()
@click.argument('model-type', type=EnumType(ModelType, False))
@click.option('-c', '--config', 'config_filepath', default=None,
              help='The path to the model configuration file. If unspecified, uses the default config for the model.')
def summary(model_type, config_filepath):
    '''
    Prints a summary of the model.

    '''

    config = composer_b200.config.get(config_filepath or get_default_config())
    model, _ = create_model(model_type, config)
    for line in model.summary_lines():
        click.echo(line)


@cli.command()
@click.argument('model-type', type=EnumType(ModelType, False))
@click.argument('restoredir')
@click.argument('output-filepath')
@click.option('--prompt', '-p', 'prompt', default=None, help='The path of the MIDI file (or preprocessed .data file) ' +
              'to prompt the network with.')
@click.option('--prompt-length', default=10, help='Number of events to take from the start of the prompt. Defaults to 10.')
@click.option('--length', '-l', 'generate_length', default=1024, help='The length of the generated event sequence. Defaults to 1024')
@click.option('--temperature', default=1.0, help='Dictates how random the result is. Low temperature yields more predictable output. ' +
              'On the other hand, high temperature yields very random ("surprising") outputs. Defaults to 1.0.')
@click.option('--count', default=1, help='Number of independent continuations to sample (extension). With more than ' +
              'one, files are numbered and sharded over the ranks of a torchrun launch.')
@click.pass_context
def generate(ctx, model_type, restoredir, output_filepath, prompt, prompt_length, generate_length, temperature, count):
    '''
    Generate a MIDI file.

    '''

    rank, world_size, _ = _distributed_context()
    seed = _agreed_seed(ctx)
    config = get_config_from_restoredir(restoredir)
    model, _ = create_model(model_type, config, seed=seed)
    model.load_from_checkpoint(restoredir)

    model.compile(get_learning_rate(model_type, config))
    model.build(input_shape=(1, None))

    if prompt is None:
        raise NotImplementedError()     # as in the reference (cli.py:642-643)

    settings = (config.dataset.time_step_increment, config.dataset.max_time_steps, config.dataset.velocity_bins)
    if str(prompt).endswith('.data'):
        event_sequence = EventSequence.from_file(prompt)
    else:
        prompt_note_sequence = NoteSequence.from_midi(prompt).trim_start()
        event_sequence = prompt_note_sequence.to_event_sequence(*settings)

    event_sequence.events = event_sequence.events[:prompt_length]

    def _encode(event):
        return IntegerEncodedEventSequence.event_to_id(event.type, event.value, event_sequence.event_ranges,
                                                       event_sequence.event_value_ranges)

    def _decode(event_id):
        return IntegerEncodedEventSequence.id_to_event(event_id, event_sequence.event_ranges,
                                                       event_sequence.event_value_ranges)

    from composer_b200 import parallel
    x = np.asarray([[_encode(event) for event in event_sequence.events]], dtype=np.int32)
    if x.shape[1] == 0:
        raise InvalidParameterError('The prompt contains no events.')

    # The reference's loop feeds back only the last sampled id without a cache (cli.py:663-676), i.e. it
    # discards the context; this build decodes with the model's KV cache (its `past=` semantics).
    first, last = parallel.shard_range(count, world_size, rank)
    if last > first:
        prompts = np.repeat(x, last - first, axis=0)
        model.reset_states()
        ids = model.generate(prompts, generate_length, temperature=temperature, seed=seed,
                             sequence_index_base=first).cpu().numpy()
        output_filepath = Path(output_filepath)
        output_filepath.parent.mkdir(parents=True, exist_ok=True)
        for row, index in zip(ids, range(first, last)):
            events = list(event_sequence.events) + [_decode(int(event_id)) for event_id in row]
            result = EventSequence(events, *settings)
            target = output_filepath if count == 1 else \
                output_filepath.with_name('{}-{}{}'.format(output_filepath.stem, index, output_filepath.suffix))
            result.to_note_sequence().to_midi(str(target))
            logging.info('Wrote \'{}\'.'.format(target))


def _not_provided(name):
    def command(*args, **kwargs):
        logging.error('The \'{}\' command is outside the Transformer train/generate path and is not provided by '
                      'the B200 build; use the reference implementation for it.'.format(name))
        exit(1)

    command.__name__ = name.replace('-', '_')
    command.__doc__ = 'Not provided by the B200 build (outside the Transformer hot path).'
    return command


for _name in ('preprocess', 'export-dataset', 'visualize-training', 'synthesize'):
    cli.command(name=_name, context_settings=dict(ignore_unknown_options=True, allow_extra_args=True))(
        click.argument('args', nargs=-1, type=click.UNPROCESSED)(_not_provided(_name)))
