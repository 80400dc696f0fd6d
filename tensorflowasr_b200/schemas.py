"""NamedTuples of the reference's API surface (tensorflow_asr/schemas.py:20-62), with torch.Tensor fields."""
from __future__ import annotations

import typing

import torch


class TrainInput(typing.NamedTuple):
    inputs: torch.Tensor
    inputs_length: torch.Tensor
    predictions: torch.Tensor
    predictions_length: torch.Tensor


class TrainOutput(typing.NamedTuple):
    logits: torch.Tensor
    logits_length: torch.Tensor


class TrainLabel(typing.NamedTuple):
    labels: torch.Tensor
    labels_length: torch.Tensor


class TrainData(typing.NamedTuple):
    inputs: TrainInput
    labels: TrainLabel


class PredictInput(typing.NamedTuple):
    inputs: torch.Tensor
    inputs_length: torch.Tensor
    previous_tokens: typing.Optional[torch.Tensor] = None
    previous_encoder_states: typing.Optional[torch.Tensor] = None
    previous_decoder_states: typing.Optional[torch.Tensor] = None


class PredictOutput(typing.NamedTuple):
    tokens: torch.Tensor
    next_tokens: torch.Tensor
    next_encoder_states: typing.Optional[torch.Tensor] = None
    next_decoder_states: typing.Optional[torch.Tensor] = None


class PredictOutputWithTranscript(typing.NamedTuple):
    transcript: typing.Any
    tokens: torch.Tensor
    next_tokens: torch.Tensor
    next_encoder_states: typing.Optional[torch.Tensor] = None
    next_decoder_states: typing.Optional[torch.Tensor] = None
