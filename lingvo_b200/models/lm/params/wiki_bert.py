"""Encoder-only BERT on the GShard dense builder (ref
`lingvo/tasks/lm/params/wiki_bert.py`): the MLPerf-BERT family, 2 B → 1 T parameters."""

import numpy as np
import torch

from lingvo_b200 import model_registry
from lingvo_b200.core import base_model_params
from lingvo_b200.core import gshard_builder
from lingvo_b200.core import optimizer
from lingvo_b200.core import program
from lingvo_b200.core import schedule
from lingvo_b200.models.lm import input_generator


class BertTemplate(base_model_params.SingleTaskModelParams):
  """Template (ref :30). `DEVICE_MESH_SHAPE = [data, model]` maps onto `[dp, tp]` ranks."""

  BATCH_SIZE = 64 * 8
  SEQUENCE_LENGTH = 512
  HIDDEN_DIM = 4096
  ATTENTION_KEY_VALUE_DIM = 128
  MODEL_DIM = 1024
  DROPOUT_RATE = 0.0
  ATTENTION_DROPOUT_RATE = 0.0
  NUM_HEADS = 8
  NUM_TRANSFORMER_LAYERS = 8
  LABEL_SMOOTHING = 0.0
  VOCAB_SIZE = 32000
  DEVICE_MESH_SHAPE = [64, 1]
  DEVICE_MESH = None
  MODEL_DIM_RESHAPE_SEGMENTS = None
  LOSS_DENOMINATOR = 0
  BETA1 = 0
  WARMUP_STEPS = 10000
  TRAIN_STEPS_PER_LOOP = 200
  TRAIN_EXES_PER_EVAL = 5
  POSITIONAL_EMBEDDING = False
  USE_REPEAT_LAYER = False
  GATED_FFN_ACT = 'silu'
  ATTEN_LOGIT_CAP = 0
  REMOVE_MASK = False

  def Task(self):
    b = gshard_builder.DenseBuilder.Params().Set(
        atten_logit_cap=self.ATTEN_LOGIT_CAP, attention_num_memory_heads=1,
        device_mesh_shape=self.DEVICE_MESH_SHAPE, device_mesh=self.DEVICE_MESH,
        relative_attention_num_buckets=32, relative_attention_type='bias',
        relative_attention_max_distance=128, dtype=torch.float32,
        fprop_dtype=torch.bfloat16, attention_logits_dtype=torch.float32,
        dropout_rate=self.DROPOUT_RATE, num_devices=1,
        attention_dropout_prob=self.ATTENTION_DROPOUT_RATE,
        attention_key_value_dim=self.ATTENTION_KEY_VALUE_DIM, attention_extra_logit=None,
        relative_attention_use_universal_1d_position=True,
        model_dim_reshape_segments=self.MODEL_DIM_RESHAPE_SEGMENTS, emb_w_split=[1, 0],
        kv_mhd_w_split=[1, -1, -1], emb_out_split=[0, -1, 1], blm_split=[0, -1, 1],
        logits_split=[0, -1, 1], model_dim=self.MODEL_DIM,
        attention_num_heads=self.NUM_HEADS, ff_dim=self.HIDDEN_DIM,
        attention_combine_dims=True)
    p = gshard_builder.BertTransformer.Params().Set(
        name='transformer', builder=b, use_repeat_layer=self.USE_REPEAT_LAYER,
        gated_ffn_activation=self.GATED_FFN_ACT,
        positional_embedding=self.POSITIONAL_EMBEDDING, dtype=torch.float32,
        fprop_dtype=torch.bfloat16, batch_size=self.BATCH_SIZE,
        sequence_length=self.SEQUENCE_LENGTH,
        num_transformer_layers=self.NUM_TRANSFORMER_LAYERS, aux_loss_coef=0.0,
        loss_denominator=self.LOSS_DENOMINATOR, label_smoothing=self.LABEL_SMOOTHING,
        vocab_size=self.VOCAB_SIZE, max_length=self.SEQUENCE_LENGTH)
    p.train.optimizer = optimizer.XLAShardingAdafactor.Params().Set(
        beta1=self.BETA1, beta2=0.99, multiply_by_parameter_scale=True,
        clipping_threshold=1.0, factored=True, decay_exponent_pow=0.8)
    p.train.learning_rate = 1.0
    p.train.lr_schedule = schedule.SqrtDecay.Params().Set(
        warmup_steps=self.WARMUP_STEPS, multiplier=1.0)
    p.train.Set(max_steps=10000000, save_max_to_keep=40,
                save_keep_checkpoint_every_n_hours=12)
    return p


class MLPerfTrainTemplate(BertTemplate):
  """MLPerf BERT data + program schedule (ref :133)."""

  TRAIN_DATA = 'gs://mlperf_v1_1/bert/train'
  EVAL_DATA = 'gs://mlperf_v1_1/bert/eval'

  def Task(self):
    p = super().Task()
    p.mask_token_id = 103
    p.masked_lm.mask_token_id = 103
    return p

  def Train(self):
    return input_generator.TFRecordBertInput.Params().Set(
        name='train', resettable=True, batch_size=self.BATCH_SIZE, enable_packing=True,
        shuffle=True, input_file=self.TRAIN_DATA, remove_mask=self.REMOVE_MASK)

  def Test(self):
    return input_generator.TFRecordBertInput.Params().Set(
        name='test', input_file=self.EVAL_DATA, batch_size=512)

  def ProgramSchedule(self):
    p = program.SimpleProgramScheduleForTask(
        train_dataset_name='Train', train_steps_per_loop=self.TRAIN_STEPS_PER_LOOP,
        eval_dataset_names=['Test'], eval_steps_per_loop=10, decode_steps_per_loop=0)
    p.train_executions_per_eval = self.TRAIN_EXES_PER_EVAL
    if p.ml_perf is not None:
      p.ml_perf.Set(benchmark_name='bert', decoder_metric_name='acc1',
                    decoder_metric_success_threshold=0.6, max_steps_to_train=31790,
                    steps_per_epoch=1 / self.BATCH_SIZE, global_batch_size=self.BATCH_SIZE,
                    max_sequence_length=self.SEQUENCE_LENGTH)
    return p


def _Mesh(shape):
  return np.arange(0, int(np.prod(shape))).reshape(shape)


@model_registry.RegisterSingleTaskModel
class MLPerfTrainBertDense2B(MLPerfTrainTemplate):
  """2 B parameters (ref :195)."""
  VOCAB_SIZE = 30522
  BATCH_SIZE = 1024
  USE_REPEAT_LAYER = True
  NUM_TRANSFORMER_LAYERS = 8
  MODEL_DIM = 4096
  NUM_HEADS = 16
  HIDDEN_DIM = 16384
  ATTENTION_KEY_VALUE_DIM = 256
  DEVICE_MESH_SHAPE = [16, 4]
  DEVICE_MESH = _Mesh(DEVICE_MESH_SHAPE)
  MODEL_DIM_RESHAPE_SEGMENTS = [4]


@model_registry.RegisterSingleTaskModel
class MLPerfBertDense1T(MLPerfTrainTemplate):
  """1 T parameters on 1024 devices (ref :214)."""
  BATCH_SIZE = 1024
  USE_REPEAT_LAYER = True
  NUM_TRANSFORMER_LAYERS = 128
  HIDDEN_DIM = 131072
  MODEL_DIM = 16384
  NUM_HEADS = 256
  DEVICE_MESH_SHAPE = [64, 16]
  DEVICE_MESH = _Mesh(DEVICE_MESH_SHAPE)
  HIDDEN_DIM_RESHAPE_SEGMENTS = 16
  MODEL_DIM_RESHAPE_SEGMENTS = [16, 4]


@model_registry.RegisterSingleTaskModel
class MLPerfBertDense1TWider(MLPerfBertDense1T):
  """1 T parameters, fewer and wider layers (ref :232)."""
  BATCH_SIZE = 4096
  NUM_TRANSFORMER_LAYERS = 32
  HIDDEN_DIM = 131072 * 2
  MODEL_DIM = 16384 * 2


@model_registry.RegisterSingleTaskModel
class MLPerfBertDense175B(MLPerfBertDense1T):
  """175 B parameters (ref :241)."""
  BATCH_SIZE = 1024
  HIDDEN_DIM = 12288 * 4
  ATTENTION_KEY_VALUE_DIM = 128
  MODEL_DIM = 12288
  NUM_HEADS = 96
  NUM_TRANSFORMER_LAYERS = 96
  POSITIONAL_EMBEDDING = True
  TRAIN_STEPS_PER_LOOP = 20


@model_registry.RegisterSingleTaskModel
class MLPerfBertDense500B(MLPerfBertDense1T):
  """481 B parameters (ref :254)."""
  VOCAB_SIZE = 30522
  BATCH_SIZE = 4096
  NUM_TRANSFORMER_LAYERS = 64
  LABEL_SMOOTHING = 0.1
  POSITIONAL_EMBEDDING = True
  REMOVE_MASK = True
  TRAIN_STEPS_PER_LOOP = 100
  TRAIN_EXES_PER_EVAL = 1


@model_registry.RegisterSingleTaskModel
class MLPerfBertDense500B2K(MLPerfBertDense500B):
  """481 B parameters on 2048 devices (ref :268)."""
  DEVICE_MESH_SHAPE = [256, 8]
  DEVICE_MESH = (np.arange(0, int(np.prod(DEVICE_MESH_SHAPE))).reshape([8, 16, 16])
                 .transpose([1, 2, 0]).reshape(DEVICE_MESH_SHAPE))
  HIDDEN_DIM_RESHAPE_SEGMENTS = 8
  MODEL_DIM_RESHAPE_SEGMENTS = [8]


@model_registry.RegisterSingleTaskModel
class MLPerfBertDense13B32x32(MLPerfBertDense1T):
  """13 B parameters (ref :283)."""
  BATCH_SIZE = 4096
  HIDDEN_DIM = 5120 * 4
  MODEL_DIM = 5120
  ATTENTION_KEY_VALUE_DIM = 128
  NUM_HEADS = 40
  NUM_TRANSFORMER_LAYERS = 40
  TRAIN_EXES_PER_EVAL = 1
  POSITIONAL_EMBEDDING = True
  LABEL_SMOOTHING = 0.1
  USE_REPEAT_LAYER = True
  REMOVE_MASK = True
  TRAIN_STEPS_PER_LOOP = 100
  DEVICE_MESH_SHAPE = [64, 32]
  DEVICE_MESH = np.reshape(np.arange(0, int(np.prod(DEVICE_MESH_SHAPE))), [32, 64]).transpose()
  HIDDEN_DIM_RESHAPE_SEGMENTS = 8
  MODEL_DIM_RESHAPE_SEGMENTS = [8]


@model_registry.RegisterSingleTaskModel
class BertDenseTiny(MLPerfTrainTemplate):
  """Unit-test sized member of the family."""
  VOCAB_SIZE = 128
  BATCH_SIZE = 4
  SEQUENCE_LENGTH = 32
  NUM_TRANSFORMER_LAYERS = 2
  MODEL_DIM = 32
  NUM_HEADS = 4
  HIDDEN_DIM = 64
  ATTENTION_KEY_VALUE_DIM = 8
  DEVICE_MESH_SHAPE = [1, 1]
