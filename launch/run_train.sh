#!/bin/bash
# run_train.sh of the reference on the B200-native library.  NGPU=8 launch/run_train.sh  (default: all visible GPUs, one process each).
set -e
cd "$(dirname "$0")/.."
NGPU=${NGPU:-$(nvidia-smi -L | wc -l)}
RUN="python -u -m"
if [ "$NGPU" -gt 1 ]; then RUN="python -u -m torch.distributed.run --nnodes=1 --nproc-per-node $NGPU --master-addr 127.0.0.1 --master-port ${PORT:-29511} -m"; fi
mkdir -p ./model_HLSTM_TeaStud_every10_train
# Training of Dynamic Teacher and Student on Train Data (batch 256 per GPU, gradients averaged over NCCL):
time $RUN efficientvideoclassification_youtube8m_b200.train --train_data_pattern "./yt8m/train*.tfrecord" --train_dir ./model_HLSTM_TeaStud_every10_train/ --frame_features True --feature_names "rgb, audio" --feature_sizes "1024, 128" --model "HierarchicalLstmModel" --num_inputs_to_lstm 20 --lstm_layers 2 --every_n 10 --batch_size 256 --start_new_model True --num_epochs 1 "$@" 2>&1 | tee output_HLSTM_TeaStud_every10_after_1epc
