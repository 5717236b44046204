#!/bin/bash
# run_convert_model.sh of the reference: keep the 11 model_student/* variables of the joint checkpoint.
set -e
cd "$(dirname "$0")/.."
time python -u -m efficientvideoclassification_youtube8m_b200.train_convert_model --train_dir ./model_HLSTM_TeaStud_every10_train/ --output_dir ./model_HLSTM_TeaStud_every10_finetune/ "$@" 2>&1 | tee output_HLSTM_TeaStud_every10_convertModel
