"""Batched replacements of the reference's test-time drivers:

    predict()            seq2seq/predict.py:57-128   greedy decoding, one example at a time in the reference
    evaluate()           seq2seq/evaluate.py:10-24   mean token accuracy / exact match / aux accuracy
    predict_and_save()   seq2seq/predict.py:16-54    predictions JSON consumed by GroundedScan.error_analysis

Same call signatures, same yielded tuples and the same JSON schema, but the data iterator may deliver
batches of ANY size (the reference refuses test_batch_size > 1, __main__.py:119-120): every batch is decoded
by one launch of the cluster-resident greedy kernel (``Model.greedy_decode`` -> ``gscan_greedy_decode``)
and then unpacked on the host into the per-example records the reference produces.  The per-example
semantics are those of the reference loop: at most ``max_decoding_steps + 1`` tokens (its ``<=``), a
trailing EOS is dropped together with its attention rows, the auxiliary scores sum the visual attention
over the generated steps only.
"""
from __future__ import annotations

import json
import logging
import time
from typing import Iterator, List, Optional, Tuple

import numpy as np
import torch

logger = logging.getLogger(__name__)


def sequence_accuracy(prediction: List[int], target: List[int]) -> float:
    """Position-wise token accuracy in percent (reference helpers.py:44-64): the shorter list is padded
    (prediction with 0, target with -1, so padding never matches); exact match <=> 100.0."""
    n = max(len(prediction), len(target))
    if n == 0:
        return 0.0
    pred = list(prediction) + [0] * (n - len(prediction))
    tgt = list(target) + [-1] * (n - len(target))
    return 100.0 * sum(1 for p, t in zip(pred, tgt) if p == t) / n


def _as_int_list(lengths) -> List[int]:
    if torch.is_tensor(lengths):
        return [int(x) for x in lengths.reshape(-1).tolist()]
    return [int(x) for x in np.asarray(lengths).reshape(-1).tolist()]


def predict(data_iterator: Iterator, model, max_decoding_steps: int, pad_idx: int, sos_idx: int, eos_idx: int,
            max_examples_to_evaluate: Optional[int] = None, return_attention: bool = True) -> Iterator[Tuple]:
    """Greedy-decode everything ``data_iterator`` yields (9-tuples of gSCAN_dataset.py:229-231, any batch
    size) and yield, per EXAMPLE, the 8-tuple of the reference's ``predict``:

        (input_sequence [1, len], derivation_spec [1], situation_spec [1], output_sequence list[int],
         target_sequence [1, len], attention_weights_commands, attention_weights_situations, aux_accuracy)

    with the attention lists shaped as the reference's ``.tolist()`` of its batch-1 tensors
    (steps x [1][Ti_b] and steps x [1][G*G]); ``return_attention=False`` leaves them empty (evaluate())."""
    model.eval()
    start_time = time.time()
    done = 0
    for (input_sequence, input_lengths, derivation_spec, situation, situation_spec, target_sequence,
         target_lengths, agent_positions, target_positions) in data_iterator:
        if max_examples_to_evaluate and done >= max_examples_to_evaluate:
            break
        with torch.no_grad():
            out = model.greedy_decode(input_sequence, input_lengths, situation, max_decoding_steps, sos_idx, eos_idx,
                                      return_attention=return_attention)
        # one device -> host transfer per batch (the reference pays 3 syncs per decoded token)
        tokens = out["tokens"].cpu().numpy()
        lengths = out["lengths"].cpu().numpy()
        alphas = out["attention_weights_commands"].cpu().numpy() if return_attention else None
        betas = out["attention_weights_situations"].cpu().numpy() if return_attention else None
        aux_pred = out["aux_logp"].argmax(dim=1).cpu().numpy() if out.get("aux_logp") is not None else None
        positions = target_positions.cpu().numpy() if torch.is_tensor(target_positions) else np.asarray(target_positions)
        cmd_len, tgt_len = _as_int_list(input_lengths), _as_int_list(target_lengths)
        for b in range(tokens.shape[0]):
            if max_examples_to_evaluate and done >= max_examples_to_evaluate:
                break
            n = int(lengths[b])
            output_sequence = [int(t) for t in tokens[b, :n]]
            if return_attention:
                att_cmd = [[alphas[b, s, :cmd_len[b]].tolist()] for s in range(n)]
                att_sit = [[betas[b, s].tolist()] for s in range(n)]
            else:
                att_cmd, att_sit = [], []
            aux_acc = 100.0 * float(aux_pred[b] == positions[b]) if aux_pred is not None else 0
            done += 1
            yield (input_sequence[b:b + 1, :cmd_len[b]], [derivation_spec[b]] if derivation_spec is not None else None,
                   [situation_spec[b]] if situation_spec is not None else None, output_sequence,
                   target_sequence[b:b + 1, :tgt_len[b]], att_cmd, att_sit, aux_acc)
    logger.info("Predicted for {} examples.".format(done))
    logger.info("Done predicting in {} seconds.".format(time.time() - start_time))


def evaluate(data_iterator: Iterator, model, max_decoding_steps: int, pad_idx: int, sos_idx: int, eos_idx: int,
             max_examples_to_evaluate: Optional[int] = None) -> Tuple[float, float, float]:
    """(mean token accuracy, exact match %, mean auxiliary accuracy) as seq2seq/evaluate.py:10-24."""
    accuracies, target_accuracies, exact_match = [], [], 0
    for _, _, _, output_sequence, target_sequence, _, _, aux_acc_target in predict(
            data_iterator=data_iterator, model=model, max_decoding_steps=max_decoding_steps, pad_idx=pad_idx,
            sos_idx=sos_idx, eos_idx=eos_idx, max_examples_to_evaluate=max_examples_to_evaluate,
            return_attention=False):
        accuracy = sequence_accuracy(output_sequence, target_sequence[0].tolist()[1:-1])
        if accuracy == 100:
            exact_match += 1
        accuracies.append(accuracy)
        target_accuracies.append(aux_acc_target)
    if not accuracies:
        raise ValueError("evaluate(): the data iterator yielded no examples")
    return (float(np.mean(np.array(accuracies))), (exact_match / len(accuracies)) * 100,
            float(np.mean(np.array(target_accuracies))))


def predict_and_save(dataset, model, output_file_path: str, max_decoding_steps: int, max_testing_examples=None,
                     batch_size: int = 200, **kwargs) -> str:
    """Predict all of ``dataset`` and write the predictions JSON (schema of predict.py:44-51:
    input, prediction, derivation, target, situation, attention_weights_input,
    attention_weights_situation, accuracy, exact_match, position_accuracy)."""
    output = []
    for (input_sequence, derivation_spec, situation_spec, output_sequence, target_sequence,
         attention_weights_commands, attention_weights_situations, position_accuracy) in predict(
            dataset.get_data_iterator(batch_size=batch_size), model=model, max_decoding_steps=max_decoding_steps,
            pad_idx=dataset.target_vocabulary.pad_idx, sos_idx=dataset.target_vocabulary.sos_idx,
            eos_idx=dataset.target_vocabulary.eos_idx, max_examples_to_evaluate=max_testing_examples):
        accuracy = sequence_accuracy(output_sequence, target_sequence[0].tolist()[1:-1])
        input_str_sequence = dataset.array_to_sentence(input_sequence[0].tolist(), vocabulary="input")[1:-1]
        target_str_sequence = dataset.array_to_sentence(target_sequence[0].tolist(), vocabulary="target")[1:-1]
        output_str_sequence = dataset.array_to_sentence(output_sequence, vocabulary="target")
        output.append({"input": input_str_sequence, "prediction": output_str_sequence,
                       "derivation": derivation_spec, "target": target_str_sequence, "situation": situation_spec,
                       "attention_weights_input": attention_weights_commands,
                       "attention_weights_situation": attention_weights_situations,
                       "accuracy": accuracy, "exact_match": True if accuracy == 100 else False,
                       "position_accuracy": position_accuracy})
    with open(output_file_path, mode="w") as outfile:
        json.dump(output, outfile, indent=4)
    logger.info("Wrote predictions for {} examples.".format(len(output)))
    return output_file_path
