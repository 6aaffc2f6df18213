# python run.py config.json   (the reference's entry script, run.py)
import vapc_b200

if __name__ == "__main__":
    vapc_b200.main()
