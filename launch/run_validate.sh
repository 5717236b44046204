#!/bin/bash
# run_validate.sh of the reference on the B200-native library.  NGPU=8 launch/run_validate.sh  (default: all visible GPUs, one process each).
set -e
cd "$(dirname "$0")/.."
NGPU=${NGPU:-$(nvidia-smi -L | wc -l)}
RUN="python -u -m"
if [ "$NGPU" -gt 1 ]; then RUN="python -u -m torch.distributed.run --nnodes=1 --nproc-per-node $NGPU --master-addr 127.0.0.1 --master-port ${PORT:-29511} -m"; fi
# Evaluate teacher+student after joint training (student predictions, label loss, state-matching loss)
time $RUN efficientvideoclassification_youtube8m_b200.validate --eval_data_pattern "./yt8m/validate*.tfrecord" --train_dir ./model_HLSTM_TeaStud_every10_train/ --frame_features True --feature_names "rgb, audio" --feature_sizes "1024, 128" --model "HierarchicalLstmModel" --num_inputs_to_lstm 20 --lstm_layers 2 --every_n 10 --batch_size 512 --top_k 20 --run_once True "$@" 2>&1 | tee validate_HLSTM_TeaStud_every10_train
