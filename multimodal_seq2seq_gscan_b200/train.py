"""Training driver with the call signature, log lines and checkpoint behaviour of the reference's
``seq2seq/train.py:14-149``, on the B200 path:

* data: the gym-free loader (``dataset.py``) instead of GroundedScan + per-example tensors;
* step: ``FusedTrainer.train_step`` (forward, loss, backward, one fused Adam launch with the LambdaLR factor
  folded in) instead of autograd + ``torch.optim.Adam`` + ``LambdaLR`` - the optimizer state it saves has
  ``torch.optim.Adam``'s layout, so checkpoints interoperate with the reference;
* in-training evaluation: the batched greedy decoder (``predict.evaluate``) at ``test_batch_size`` examples per
  launch - the reference decodes one example at a time and refuses ``test_batch_size > 1``;
* data parallelism (new): under ``torchrun`` every rank builds the same shuffled order (numpy seeded with
  ``seed``; the reference leaves it unseeded, train.py:27 vs gSCAN_dataset.py:179) and takes its contiguous
  shard of every global batch; gradients and loss normalisers are summed by ONE all-reduce per step (``dp.py``).
"""
from __future__ import annotations

import logging
import os
import random

import numpy as np
import torch
import torch.distributed as dist

from . import dp
from .dataset import GroundedScanDataset
from .model import Model
from .predict import evaluate
from .trainer import FusedTrainer

logger = logging.getLogger(__name__)


def train(data_path: str, data_directory: str, generate_vocabularies: bool, input_vocab_path: str,
          target_vocab_path: str, embedding_dimension: int, num_encoder_layers: int, encoder_dropout_p: float,
          encoder_bidirectional: bool, training_batch_size: int, test_batch_size: int, max_decoding_steps: int,
          num_decoder_layers: int, decoder_dropout_p: float, cnn_kernel_size: int, cnn_dropout_p: float,
          cnn_hidden_num_channels: int, simple_situation_representation: bool, decoder_hidden_size: int,
          encoder_hidden_size: int, learning_rate: float, adam_beta_1: float, adam_beta_2: float, lr_decay: float,
          lr_decay_steps: int, resume_from_file: str, max_training_iterations: int, output_directory: str,
          print_every: int, evaluate_every: int, conditional_attention: bool, auxiliary_task: bool,
          weight_target_loss: float, attention_type: str, k: int, max_training_examples=None, seed=42, **kwargs):
    if not torch.cuda.is_available():
        raise RuntimeError("multimodal_seq2seq_gscan_b200.train needs a CUDA device: there is no CPU fallback")
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    rank = dist.get_rank() if distributed else 0
    world = dist.get_world_size() if distributed else 1
    device = torch.device("cuda", torch.cuda.current_device())
    cfg = locals().copy()
    for name in ("distributed", "rank", "world", "device", "kwargs"):
        cfg.pop(name)

    torch.manual_seed(seed)
    np.random.seed(seed)       # identical shuffles on every rank
    random.seed(seed)          # ... and the same k-shot examples moved into train / dev (dataset.py uses random.sample)

    logger.info("Loading Training set...")
    training_set = GroundedScanDataset(data_path, data_directory, split="train",
                                       input_vocabulary_file=input_vocab_path,
                                       target_vocabulary_file=target_vocab_path,
                                       generate_vocabulary=generate_vocabularies, k=k, device=device)
    training_set.read_dataset(max_examples=max_training_examples,
                              simple_situation_representation=simple_situation_representation)
    logger.info("Done Loading Training set.")
    logger.info("  Loaded {} training examples.".format(training_set.num_examples))
    logger.info("  Input vocabulary size training set: {}".format(training_set.input_vocabulary_size))
    logger.info("  Most common input words: {}".format(training_set.input_vocabulary.most_common(5)))
    logger.info("  Output vocabulary size training set: {}".format(training_set.target_vocabulary_size))
    logger.info("  Most common target words: {}".format(training_set.target_vocabulary.most_common(5)))
    if generate_vocabularies and rank == 0:
        training_set.save_vocabularies(input_vocab_path, target_vocab_path)
        logger.info("Saved vocabularies to {} for input and {} for target.".format(input_vocab_path, target_vocab_path))
    if distributed:
        dist.barrier()

    logger.info("Loading Dev. set...")
    test_set = GroundedScanDataset(data_path, data_directory, split="dev", input_vocabulary_file=input_vocab_path,
                                   target_vocabulary_file=target_vocab_path, generate_vocabulary=False, k=0,
                                   device=device)
    test_set.read_dataset(max_examples=None, simple_situation_representation=simple_situation_representation)
    test_set.shuffle_data()
    logger.info("Done Loading Dev. set.")

    model = Model(input_vocabulary_size=training_set.input_vocabulary_size,
                  target_vocabulary_size=training_set.target_vocabulary_size,
                  num_cnn_channels=training_set.image_channels,
                  input_padding_idx=training_set.input_vocabulary.pad_idx,
                  target_pad_idx=training_set.target_vocabulary.pad_idx,
                  target_eos_idx=training_set.target_vocabulary.eos_idx,
                  **cfg).to(device)
    trainer = FusedTrainer(model, learning_rate=learning_rate, adam_beta_1=adam_beta_1, adam_beta_2=adam_beta_2,
                           lr_decay=lr_decay, lr_decay_steps=lr_decay_steps, weight_target_loss=weight_target_loss,
                           distributed=distributed)

    start_iteration = 1
    best_exact_match = 0
    if resume_from_file:
        assert os.path.isfile(resume_from_file), "No checkpoint found at {}".format(resume_from_file)
        logger.info("Loading checkpoint from file at '{}'".format(resume_from_file))
        optimizer_state_dict = model.load_model(resume_from_file)
        trainer.load_state_dict(optimizer_state_dict)     # (the parameters were copied in place into the flat buffer)
        start_iteration = model.trained_iterations
        best_exact_match = model.best_exact_match or 0
        logger.info("Loaded checkpoint '{}' (iter {})".format(resume_from_file, start_iteration))

    logger.info("Training starts..")
    training_iteration = start_iteration
    last_loss = None
    while training_iteration < max_training_iterations:
        training_set.shuffle_data()
        for (input_batch, input_lengths, _, situation_batch, _, target_batch,
             target_lengths, agent_positions, target_positions) in training_set.get_data_iterator(
                batch_size=training_batch_size):
            global_counts = None
            if distributed:
                # this rank's contiguous shard of the global batch (every rank holds the same global batch, so the
                # global loss normalisers are known without communication); a batch with fewer examples than ranks -
                # the ragged tail of an epoch - is skipped by EVERY rank, so nobody waits in a collective alone
                sharded = dp.shard_batch((input_batch, input_lengths, None, situation_batch, None, target_batch,
                                          target_lengths, agent_positions, target_positions), rank, world,
                                         auxiliary_task)
                if sharded is None:
                    continue
                (input_batch, input_lengths, _, situation_batch, _, target_batch, target_lengths, agent_positions,
                 target_positions), global_counts = sharded
            is_best = False
            loss = trainer.train_step(input_batch, input_lengths, situation_batch, target_batch, target_lengths,
                                      target_positions if auxiliary_task else None, global_counts=global_counts)
            last_loss = loss

            if training_iteration % print_every == 0:
                accuracy, exact_match = model.get_metrics(trainer.last_logp, target_batch)
                if auxiliary_task:
                    auxiliary_accuracy_target = model.get_auxiliary_accuracy(trainer.last_aux, target_positions)
                else:
                    auxiliary_accuracy_target = 0.
                logger.info("Iteration %08d, loss %8.4f, accuracy %5.2f, exact match %5.2f, learning_rate %.5f,"
                            " aux. accuracy target pos %5.2f" % (training_iteration, float(loss), accuracy, exact_match,
                                                                 trainer.current_lr(), auxiliary_accuracy_target))

            if training_iteration % evaluate_every == 0 and rank == 0:
                with torch.no_grad():
                    model.eval()
                    logger.info("Evaluating..")
                    accuracy, exact_match, target_accuracy = evaluate(
                        test_set.get_data_iterator(batch_size=max(1, int(test_batch_size))), model=model,
                        max_decoding_steps=max_decoding_steps, pad_idx=test_set.target_vocabulary.pad_idx,
                        sos_idx=test_set.target_vocabulary.sos_idx, eos_idx=test_set.target_vocabulary.eos_idx,
                        max_examples_to_evaluate=kwargs.get("max_testing_examples"))
                    logger.info("  Evaluation Accuracy: %5.2f Exact Match: %5.2f "
                                " Target Accuracy: %5.2f" % (accuracy, exact_match, target_accuracy))
                    if exact_match > best_exact_match:
                        is_best = True
                        best_exact_match = exact_match
                        model.update_state(accuracy=accuracy, exact_match=exact_match, is_best=is_best)
                    if is_best:
                        model.save_checkpoint(file_name="checkpoint.pth.tar", is_best=is_best,
                                              optimizer_state_dict=trainer.state_dict())

            training_iteration += 1
            if training_iteration > max_training_iterations:
                break
    logger.info("Finished training.")
    return {"iterations": training_iteration - 1, "last_loss": None if last_loss is None else float(last_loss),
            "best_exact_match": best_exact_match, "model": model, "trainer": trainer}
