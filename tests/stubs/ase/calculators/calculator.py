all_changes = ["positions", "numbers", "cell", "pbc", "initial_charges", "initial_magmoms"]


class Calculator:
    implemented_properties = ()

    def __init__(self, **kwargs):
        self.results = {}
        self.atoms = None
        self.parameters = dict(kwargs)

    def calculate(self, atoms=None, properties=None, system_changes=None):
        self.atoms = atoms
