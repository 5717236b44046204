"""`python -m efficientvideoclassification_youtube8m_b200.validate --flag value ...`: validate.py main (run_validate.sh); see launchers.validate_main."""
from .launchers import validate_main as main

if __name__ == "__main__":
    main()
