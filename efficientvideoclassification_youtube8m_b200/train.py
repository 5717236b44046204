"""`python -m efficientvideoclassification_youtube8m_b200.train --flag value ...`: train.py:704-737 (run_train.sh); see launchers.train_main."""
from .launchers import train_main as main

if __name__ == "__main__":
    main()
