class footnote:      # literature.py only isinstance()-tests against it
    pass
