"""`python -m efficientvideoclassification_youtube8m_b200.train_finetune --flag value ...`: train_finetune.py main (run_finetune.sh); see launchers.finetune_main."""
from .launchers import finetune_main as main

if __name__ == "__main__":
    main()
