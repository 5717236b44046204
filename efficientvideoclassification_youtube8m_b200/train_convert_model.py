"""`python -m efficientvideoclassification_youtube8m_b200.train_convert_model --flag value ...`: train_convert_model.py main (run_convert_model.sh); see launchers.convert_main."""
from .launchers import convert_main as main

if __name__ == "__main__":
    main()
