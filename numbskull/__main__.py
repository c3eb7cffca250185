"""``python -m numbskull`` -- the reference's command line, B200-native underneath."""
from numbskull_b200.numbskull import main

main()
