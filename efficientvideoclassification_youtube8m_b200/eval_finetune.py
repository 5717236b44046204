"""`python -m efficientvideoclassification_youtube8m_b200.eval_finetune --flag value ...`: eval_finetune.py main (run_eval.sh); see launchers.eval_main."""
from .launchers import eval_main as main

if __name__ == "__main__":
    main()
