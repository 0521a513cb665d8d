// placeholder: flow()/matvec layer (filled in next)
