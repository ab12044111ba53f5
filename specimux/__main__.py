from specimux_b200.cli import main

main()
