"""Pure-Python reader for the HDF5 files OpenFermion writes.

The reference stores molecules with h5py (`MolecularData.save`, chem/molecular_data.py:705-850)
and reads them back in `MolecularData.load` (:852-915), `get_from_file` (:917-935) and
`load_operator`-style helpers.  h5py is not installed on the GPU image, so the integrals that
feed the GPU path (config 1 of BASELINE.json: LiH from the bundled HDF5) are read here with
nothing but `struct`, `zlib` and numpy.

Supported subset (what h5py with default settings produces for those files): version-0
superblock, version-1 object headers with continuation blocks, old-style groups (symbol
table message -> v1 B-tree + local heap -> symbol nodes), datasets with compact, contiguous
or chunked (v1 chunk B-tree) layout, deflate / shuffle / fletcher32 filters, and the
datatypes fixed-point, floating-point, fixed-length string and the int8 enum h5py uses for
numpy.bool_.  Anything else raises Hdf5FormatError naming the unsupported feature.
"""
import struct
import zlib

import numpy

UNDEFINED = 0xFFFFFFFFFFFFFFFF


class Hdf5FormatError(ValueError):
    """The file uses an HDF5 feature outside the supported subset (or is corrupt)."""


class _Datatype:
    def __init__(self, dtype, size, is_bool=False):
        self.dtype = dtype
        self.size = size
        self.is_bool = is_bool


class Dataset:
    """One dataset: `shape`, `dtype` and `read()` (the `[...]` of h5py)."""

    def __init__(self, file, name, shape, datatype, layout, filters):
        self._file = file
        self.name = name
        self.shape = shape
        self._datatype = datatype
        self._layout = layout
        self._filters = filters
        self.dtype = numpy.dtype(bool) if datatype.is_bool else datatype.dtype

    def read(self):
        raw = self._file._read_layout(self._layout, self.shape, self._datatype.size, self._filters)
        array = numpy.frombuffer(raw, dtype=self._datatype.dtype,
                                 count=int(numpy.prod(self.shape, dtype=numpy.int64))).reshape(self.shape)
        if self._datatype.is_bool:
            return array.astype(bool)
        return array.copy()

    def __getitem__(self, key):
        if key is not Ellipsis and key != ():
            return self.read()[key]
        return self.read()


class Group:
    def __init__(self, file, name, entries):
        self._file = file
        self.name = name
        self._entries = entries  # name -> object header address

    def keys(self):
        return list(self._entries)

    def __contains__(self, name):
        try:
            self[name]
            return True
        except KeyError:
            return False

    def __getitem__(self, path):
        node = self
        for part in [p for p in path.split('/') if p]:
            if not isinstance(node, Group) or part not in node._entries:
                raise KeyError(path)
            node = node._file._object(node._entries[part], node.name.rstrip('/') + '/' + part)
        return node


class File(Group):
    """`File(path)[name][...]` mirrors the h5py calls of chem/molecular_data.py:852-935."""

    def __init__(self, path):
        with open(path, 'rb') as handle:
            self._raw = handle.read()
        raw = self._raw
        if raw[:8] != b'\x89HDF\r\n\x1a\n':
            raise Hdf5FormatError('not an HDF5 file (signature at offset 0 missing)')
        if raw[8] != 0:
            raise Hdf5FormatError('superblock version {} is not supported (only 0)'.format(raw[8]))
        if raw[13] != 8 or raw[14] != 8:
            raise Hdf5FormatError('only 8-byte offsets and lengths are supported')
        self._base = struct.unpack_from('<Q', raw, 24)[0]
        header_address = struct.unpack_from('<Q', raw, 64)[0]
        root = self._object(header_address, '/')
        if not isinstance(root, Group):
            raise Hdf5FormatError('root object is not a group')
        super().__init__(self, '/', root._entries)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    # ---- low level -----------------------------------------------------------------------
    def _u(self, fmt, offset):
        return struct.unpack_from('<' + fmt, self._raw, offset)

    def _messages(self, address):
        """(type, data offset, size) of every message of a version-1 object header."""
        raw = self._raw
        address += self._base
        if raw[address] != 1:
            if raw[address:address + 4] == b'OHDR':
                raise Hdf5FormatError('version-2 object headers are not supported')
            raise Hdf5FormatError('bad object header version {}'.format(raw[address]))
        count, _, size = self._u('HII', address + 2)
        blocks = [(address + 16, size)]
        out = []
        while blocks and len(out) < count:
            start, length = blocks.pop(0)
            pos, end = start, start + length
            while pos + 8 <= end and len(out) < count:
                mtype, msize, _flags = self._u('HHB', pos)
                data = pos + 8
                if mtype == 0x0010:  # continuation
                    off, ln = self._u('QQ', data)
                    blocks.append((off + self._base, ln))
                out.append((mtype, data, msize))
                pos = data + msize
        return out

    def _object(self, address, name):
        messages = self._messages(address)
        kinds = {m[0] for m in messages}
        if 0x0011 in kinds:
            data = next(m[1] for m in messages if m[0] == 0x0011)
            btree, heap = self._u('QQ', data)
            return Group(self, name, self._group_entries(btree, heap))
        if 0x0002 in kinds or 0x0006 in kinds:
            raise Hdf5FormatError('new-style (link message) groups are not supported: ' + name)
        if 0x0008 in kinds:
            return self._dataset(messages, name)
        raise Hdf5FormatError('object {} is neither an old-style group nor a dataset'.format(name))

    # ---- groups --------------------------------------------------------------------------
    def _heap_string(self, heap_address, offset):
        raw = self._raw
        heap_address += self._base
        if raw[heap_address:heap_address + 4] != b'HEAP':
            raise Hdf5FormatError('local heap signature missing')
        data_address = self._u('Q', heap_address + 24)[0] + self._base
        start = data_address + offset
        end = raw.index(b'\x00', start)
        return raw[start:end].decode('utf-8')

    def _group_entries(self, btree_address, heap_address):
        entries = {}
        if btree_address == UNDEFINED:
            return entries
        raw = self._raw
        stack = [btree_address + self._base]
        while stack:
            node = stack.pop()
            if raw[node:node + 4] == b'TREE':
                node_type, _level, used = self._u('BBH', node + 4)
                if node_type != 0:
                    raise Hdf5FormatError('group B-tree node of type {}'.format(node_type))
                children = [self._u('Q', node + 24 + 8 + 16 * k)[0] + self._base for k in range(used)]
                stack.extend(reversed(children))
            elif raw[node:node + 4] == b'SNOD':
                count = self._u('H', node + 6)[0]
                for k in range(count):
                    entry = node + 8 + 40 * k
                    name_offset, header = self._u('QQ', entry)
                    entries[self._heap_string(heap_address, name_offset)] = header
            else:
                raise Hdf5FormatError('unexpected node signature in a group B-tree')
        return entries

    # ---- datasets ------------------------------------------------------------------------
    def _datatype(self, pos):
        class_and_version, b0, b1, _b2, size = self._u('BBBBI', pos)
        klass = class_and_version & 0x0F
        order = '>' if (b0 & 1) else '<'
        if klass == 0:  # fixed point
            signed = bool(b0 & 0x08)
            return _Datatype(numpy.dtype('{}{}{}'.format(order, 'i' if signed else 'u', size)), size), pos + 8 + 4
        if klass == 1:  # floating point
            if size not in (2, 4, 8):
                raise Hdf5FormatError('floating-point size {} is not supported'.format(size))
            return _Datatype(numpy.dtype('{}f{}'.format(order, size)), size), pos + 8 + 12
        if klass == 3:  # fixed-length string
            return _Datatype(numpy.dtype('S{}'.format(size)), size), pos + 8
        if klass == 8:  # enumeration: h5py's numpy.bool_ is an int8 enum FALSE=0, TRUE=1
            members = b0 | (b1 << 8)
            base, after = self._datatype(pos + 8)
            names = []
            version = class_and_version >> 4
            for _ in range(members):
                end = self._raw.index(b'\x00', after)
                names.append(self._raw[after:end])
                length = end - after + 1
                after += length if version >= 3 else (length + 7) // 8 * 8
            after += members * base.size
            is_bool = sorted(names) == [b'FALSE', b'TRUE'] and base.size == 1
            return _Datatype(base.dtype, base.size, is_bool), after
        names = {2: 'time', 4: 'bitfield', 5: 'opaque', 6: 'compound', 7: 'reference', 9: 'variable-length',
                 10: 'array'}
        raise Hdf5FormatError('datatype class {} is not supported'.format(names.get(klass, klass)))

    def _dataspace(self, pos):
        version, rank, flags = self._u('BBB', pos)
        if version == 1:
            dims_at = pos + 8
        elif version == 2:
            if self._raw[pos + 3] == 2:  # null dataspace
                return None
            dims_at = pos + 4
        else:
            raise Hdf5FormatError('dataspace version {}'.format(version))
        del flags
        return tuple(self._u('Q', dims_at + 8 * k)[0] for k in range(rank))

    def _filters(self, pos):
        version, count = self._u('BB', pos)
        pos += 8 if version == 1 else 2
        filters = []
        for _ in range(count):
            if version == 1:
                fid, name_len, _flags, n_values = self._u('HHHH', pos)
                pos += 8 + (name_len + 7) // 8 * 8
            else:
                fid = self._u('H', pos)[0]
                if fid >= 256:
                    name_len, _flags, n_values = self._u('HHH', pos + 2)
                    pos += 8 + name_len
                else:
                    _flags, n_values = self._u('HH', pos + 2)
                    pos += 6
            values = [self._u('I', pos + 4 * k)[0] for k in range(n_values)]
            pos += 4 * n_values
            if version == 1 and n_values % 2:
                pos += 4
            filters.append((fid, values))
        return filters

    def _dataset(self, messages, name):
        shape = datatype = layout = None
        filters = []
        for mtype, data, _size in messages:
            if mtype == 0x0001:
                shape = self._dataspace(data)
            elif mtype == 0x0003:
                datatype, _ = self._datatype(data)
            elif mtype == 0x000B:
                filters = self._filters(data)
            elif mtype == 0x0008:
                version, klass = self._u('BB', data)
                if version != 3:
                    raise Hdf5FormatError('data layout version {} is not supported'.format(version))
                if klass == 0:
                    size = self._u('H', data + 2)[0]
                    layout = ('compact', data + 4, size)
                elif klass == 1:
                    address, size = self._u('QQ', data + 2)
                    layout = ('contiguous', address, size)
                elif klass == 2:
                    rank = self._raw[data + 2]
                    btree = self._u('Q', data + 3)[0]
                    dims = [self._u('I', data + 11 + 4 * k)[0] for k in range(rank)]
                    layout = ('chunked', btree, dims)
                else:
                    raise Hdf5FormatError('data layout class {}'.format(klass))
        if shape is None or datatype is None or layout is None:
            raise Hdf5FormatError('dataset {} lacks a dataspace, datatype or layout message'.format(name))
        return Dataset(self, name, shape, datatype, layout, filters)

    def _defilter(self, chunk, filters, mask, element_size):
        for index in range(len(filters) - 1, -1, -1):
            if mask & (1 << index):
                continue
            fid, _values = filters[index]
            if fid == 1:
                chunk = zlib.decompress(chunk)
            elif fid == 2:
                count = len(chunk) // element_size
                planes = numpy.frombuffer(chunk, dtype=numpy.uint8, count=count * element_size)
                chunk = planes.reshape(element_size, count).T.tobytes() + chunk[count * element_size:]
            elif fid == 3:
                chunk = chunk[:-4]
            else:
                raise Hdf5FormatError('filter {} is not supported'.format(fid))
        return chunk

    def _read_layout(self, layout, shape, element_size, filters):
        kind = layout[0]
        total = int(numpy.prod(shape, dtype=numpy.int64)) * element_size
        if kind == 'compact':
            return self._raw[layout[1]:layout[1] + layout[2]][:total]
        if kind == 'contiguous':
            if layout[1] == UNDEFINED:
                return bytes(total)
            start = layout[1] + self._base
            return self._raw[start:start + total]
        btree, dims = layout[1], layout[2]
        chunk_shape = tuple(dims[:-1])
        rank = len(chunk_shape)
        if rank != len(shape):
            raise Hdf5FormatError('chunk rank differs from the dataspace rank')
        out = numpy.zeros(shape, dtype=numpy.dtype('V{}'.format(element_size)))
        if btree == UNDEFINED:
            return out.tobytes()
        stack = [btree + self._base]
        raw = self._raw
        key_size = 8 + 8 * (rank + 1)
        while stack:
            node = stack.pop()
            if raw[node:node + 4] != b'TREE':
                raise Hdf5FormatError('chunk B-tree signature missing')
            node_type, level, used = self._u('BBH', node + 4)
            if node_type != 1:
                raise Hdf5FormatError('chunk B-tree node of type {}'.format(node_type))
            for k in range(used):
                key = node + 24 + k * (key_size + 8)
                size, mask = self._u('II', key)
                offsets = [self._u('Q', key + 8 + 8 * d)[0] for d in range(rank)]
                child = self._u('Q', key + key_size)[0] + self._base
                if level > 0:
                    stack.append(child)
                    continue
                chunk = self._defilter(raw[child:child + size], filters, mask, element_size)
                block = numpy.frombuffer(chunk, dtype=out.dtype,
                                         count=int(numpy.prod(chunk_shape, dtype=numpy.int64))).reshape(chunk_shape)
                window = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offsets, chunk_shape, shape))
                clip = tuple(slice(0, w.stop - w.start) for w in window)
                out[window] = block[clip]
        return out.tobytes()


def load_molecular_data(path):
    """The fields `MolecularData.load` and the lazy properties read (chem/molecular_data.py:
    852-915 and get_from_file :917-935), as a plain dict.  Missing calculations are stored by
    the reference as the scalar False and come back as None, like the reference's
    `data.dtype.num != 0` test does."""
    if not path.endswith('.hdf5'):
        path = path + '.hdf5'
    out = {}
    with File(path) as f:
        atoms = f['geometry/atoms']
        if atoms.shape != ():
            out['geometry'] = [(atom.tobytes().decode('utf-8'), [float(v) for v in pos])
                               for atom, pos in zip(atoms[...], f['geometry/positions'][...])]
        else:
            out['geometry'] = atoms[...].tobytes().decode('utf-8')
        for key in ('basis', 'name'):
            out[key] = f[key][...].tobytes().decode('utf-8')
        out['description'] = f['description'][...].tobytes().decode('utf-8').rstrip('\x00')
        for key in ('multiplicity', 'charge', 'n_atoms', 'n_electrons'):
            out[key] = int(f[key][...])
        out['atoms'] = f['atoms'][...]
        out['protons'] = f['protons'][...]
        for key, cast in (('n_orbitals', int), ('n_qubits', int), ('nuclear_repulsion', float)):
            data = f[key][...]
            out[key] = cast(data) if data.dtype != bool else None
        optional = ('hf_energy', 'orbital_energies', 'mp2_energy', 'cisd_energy', 'fci_energy', 'ccsd_energy',
                    'canonical_orbitals', 'overlap_integrals', 'one_body_integrals', 'two_body_integrals',
                    'cisd_one_rdm', 'cisd_two_rdm', 'fci_one_rdm', 'fci_two_rdm', 'ccsd_single_amps',
                    'ccsd_double_amps')
        for key in optional:
            if key in f:
                data = f[key][...]
                out[key] = data if data.dtype != bool else None
            else:
                out[key] = None
    return out


def load_molecular_hamiltonian(path):
    """`MolecularData(filename=path).get_molecular_hamiltonian()` (chem/molecular_data.py:
    1011-1039) as a hamiltonians.InteractionTensor: spin-orbital expansion of the stored spatial
    integrals, two-body part halved; accepted by linalg.get_sparse_operator and
    hamiltonians.interaction_operator_jw."""
    from openfermion_b200 import hamiltonians
    data = load_molecular_data(path)
    if data['one_body_integrals'] is None or data['two_body_integrals'] is None:
        raise ValueError('the file holds no molecular integrals')
    return hamiltonians.molecular_hamiltonian(data['nuclear_repulsion'], data['one_body_integrals'],
                                        data['two_body_integrals'])
