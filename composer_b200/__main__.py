from composer_b200.cli import cli

if __name__ == '__main__':
    cli()
