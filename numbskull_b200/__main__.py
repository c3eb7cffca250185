"""``python -m numbskull_b200`` == the reference's ``numbskull`` console script."""
from .numbskull import main

main()
